/* vs_internal.h -- structures shared by the host side (vs_api.cu) and the kernels (vs_kernels.cu).
 * Not part of the ABI. */
#ifndef VS_INTERNAL_H
#define VS_INTERNAL_H
#include <stdint.h>

#define VS_ORDER   22          /* vowel_new.c:172 */
#define VS_RING    24          /* state ring / samples per unrolled filter block (>= VS_ORDER, 3 x 16 B of PCM) */
#define VS_NT      128         /* rows per CTA of the render kernel (and its RNG stride)           */
#define VS_PLAN_NT 512         /* threads per CTA of the plan kernel.  With VS_PLAN_SMEM of shared memory a plan
                                  CTA cannot share an SM with a render CTA: the plan kernel of call k+1 runs on
                                  the VS_PLAN_SMS SMs the chunk planner keeps free, instead of being starved of
                                  issue slots by the render warps of call k (measured: 0.35 -> 0.7 ms)        */
#define VS_PLAN_SMEM (96 * 1024)
#define VS_PLAN_SMS  8
/* batches up to this many streams get one WARP per stream in the plan kernel (vs_api.cu, plan launch) */
#define VS_PLAN_WARP_MAX       1024    /* while the previous call is still on the GPU */
#define VS_PLAN_WARP_MAX_IDLE  6144    /* GPU idle: the plan kernel is the call's critical path */
#define VS_PLAN_WARP_MAX_NOISE 12288   /* some streams carry glottal noise */
#define VS_RNG_DEG 31          /* glibc TYPE_3 */
#define VS_NO_CHUNK 0xffffffffu
#define VS_PULSE_TABLE_CAP (16u << 20)   /* doubles of cosine / pulse tables a context may accumulate (128 MB) */
#define VS_COS_SLACK 64         /* doubles readable past the last cosine table: a work item of the render kernel
                                  loads 64 table entries whatever the length of its segment (results discarded) */
#define VS_WIN      192        /* samples per stream per render window (8 ring blocks, 24 x 16 B): flow-only mode */
#define VS_WIN_WIDE 240        /* the same for the fused / filter-only kernels (10 ring blocks, 30 x 16 B: one
                                  16-byte piece per lane in the write-out)                                                     */

/* per-stream descriptor, prepared on the host, read once per thread */
struct VsStream {
    uint32_t n;            /* samples in the stream, flowgen_shimmer.c:242                        */
    int32_t  P;            /* nominal period (int)((float)fs/F0), :244                            */
    int32_t  T2;           /* ceil(0.5*cq*P), :317                                                */
    uint32_t cos_off;      /* first entry of this T2's cosine table                               */
    int32_t  amp;
    float    DC;
    float    jitter, shimmer, K, Kvar, noise;
    uint32_t seed;
    float    gain, pre;
    int16_t  DCs;          /* (short)DC                                                           */
    uint8_t  flags;        /* VS_F_*                                                              */
    uint8_t  preset;       /* index into "aiu1234567"                                             */
    uint32_t chunk0;       /* id of the stream's first chunk                                      */
    uint32_t n_chunks;
    uint32_t tab_cap;      /* period-table capacity                                               */
    uint32_t pulse_off;    /* first entry of the table the render kernel evaluates the open phase from: h[0..T2) then either
                              c[0..T2) (== cos_off) or, for a stream whose closure speed never varies, K*c - K + 1           */
    uint64_t out_off;      /* samples, relative to pcm_out (and raw_out)                          */
    uint64_t in_off;       /* samples, relative to flow_in (filter-only mode)                     */
    uint64_t tab_off;      /* first entry of the stream's period table                            */
    uint64_t log_off;      /* first record of the stream's user-facing period log                 */
};

/* one time-chunk of one stream = one render thread */
struct VsChunk {
    uint32_t stream;
    uint32_t emit_lo, emit_hi;   /* samples [emit_lo, emit_hi) are written by this chunk           */
    uint32_t gen_target;         /* generation starts at the period containing this sample         */
    uint32_t first_period;       /* filled by the plan kernel                                      */
    uint32_t pad[3];
};

/* period table entry (32 B) written by the plan kernel, read by the render kernel */
struct VsPeriod {
    double   Ad;           /* (double)Amplitude                                                   */
    double   Kd;           /* (double)Knew                                                        */
    uint32_t start;        /* first sample of the period                                          */
    uint32_t T_np;         /* period length T (low 16 bits) | random() draws of the period before
                              its first noise draw (high 16 bits)                                 */
    uint32_t T34;          /* T3 (low 16) | T4 (high 16): closure / DC-crossing instants (noise)  */
    int32_t  ndw;          /* NoiseDistWidth                                                      */
};

/* kernel launch argument blocks */
struct VsPlanArgs {
    const VsStream *streams;
    uint32_t        n_streams;
    VsChunk        *chunks;
    VsPeriod       *table;
    uint32_t       *rng_snap;       /* [n_chunks][32] or NULL                                     */
    uint32_t       *n_periods;      /* [n_streams]                                                */
    const double   *costab;
    void           *log;            /* vs_period_rec* or NULL                                     */
    int32_t        *status;         /* device error flag                                          */
    int             need_pulse;     /* evaluate the pulse for every stream (log requested)        */
};

struct VsRenderArgs {
    const VsStream *streams;
    const VsChunk  *chunks;
    const uint32_t *order;          /* render row t works on chunk order[t]; VS_NO_CHUNK = padding   */
    double          ncf[VS_RING];   /* -A[j] of the ONE vowel preset this launch serves: as kernel
                                       parameters they are constant-bank immediates of the DFMAs   */
    uint32_t        n_rows;         /* rows incl. padding, multiple of VS_NT                         */
    uint32_t        n_chunks;
    const VsPeriod *table;
    const uint32_t *rng_snap;
    const double   *costab;
    const double   *coef;           /* [10][24] denominators, device copy                         */
    const int16_t  *flow_in;        /* filter-only mode                                           */
    int16_t        *pcm_out;
    double         *raw_out;        /* nullable                                                   */
    int             checked_quant;  /* 1: |waveform| may reach 2^30, use the range-checked quantiser */
    int             general_pulse;  /* 1: some stream has -z Kvar > 0, its falling branch needs Knew of each period; 0: every
                                       pulse_off table already holds K*c - K + 1, a sample is ceil(A * table[i])              */
    int32_t        *status;         /* device error flag (shared with the plan kernel)              */
};

/* vowel -n (N1): one entry per stream */
struct VsNoiseRow {
    uint64_t off;          /* samples, relative to the pcm pointer */
    uint32_t n;
    uint32_t frame;        /* 50 * ((int)(fs*0.001/2.0)*2), vowel_new.c:361-363 */
    float    snr;
    uint32_t seed;
};

#endif
