/* vs_kernels.cu -- sm_100a kernels of libvoicesynth_cuda.
 *
 *   vs_plan_kernel       one THREAD per stream: the strictly sequential part of flowgen_shimmer.c:246-423
 *                        -- glibc random() state, jitter and shimmer random walks with their rejection
 *                        loops, the closure-speed draw, and (with -n) the pulse power and the number of
 *                        noise draws.  Emits a period table and, per time-chunk, the period the chunk
 *                        starts in plus the RNG state there.  The throughput form (many streams).
 *   vs_plan_warp_kernel  the same table from one WARP per stream: random() 31 values per step by
 *                        shuffles, the values' meanings worked out lane-parallel, the walk warp-uniform.
 *                        The latency form (few or long streams, glottal noise).
 *
 *   vs_render_kernel     one row per (stream, time-chunk), 128 rows per CTA, no block-level
 *                        synchronisation.  Per 32 rows ONE consumer warp and FOUR producer warps share
 *                        two shared-memory tiles [32 rows][192 or 240 int16] through named barriers:
 *                          G  generate (producers): given the period table a sample is a pure function
 *                             of its index (flowgen_shimmer.c:319,328,335); lane-parallel bookkeeping
 *                             queues per-period segments, the open phases are evaluated as flat lists
 *                             of 64-sample work items, the closed phase is a word fill.
 *                          F  filter (consumer): each lane runs the order-22 FP64 all-pole recurrence
 *                             (vowel_new.c:266-289) down its own row, in place.  The state is a
 *                             24-entry register ring, fully unrolled: ~25 FP64-pipe instructions per
 *                             sample -> the FP64 pipe is the bound.
 *                          W  write (producers): rows leave as whole 16-byte pieces, consecutive
 *                             lanes on consecutive pieces of a row (full 32-byte sectors).
 *                        Flow-only mode has no consumer: the producers do G then W.
 *   vs_vnoise_kernel     vowel -n (vowel_new.c:302-324), one thread per stream.
 *
 * Every operation that decides an integer (period length, amplitude, sample value, draw count) is
 * written with explicit round-to-nearest intrinsics in the reference's evaluation order, so those
 * results are bit-exact.  Only the filter's multiply-subtract is contracted to FMA (fast mode).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/voicesynth.h"
#include "vs_internal.h"

#define VS_RAND_MAX_D 2147483647.0

/* Filter coefficients reach the DFMAs as constant-bank operands (kernel parameters, VsRenderArgs::ncf;
 * one launch per vowel preset).  On B200 a DFMA with two register operands + one constant operand
 * issues every ~2.2 cycles per SM sub-partition, with three register operands only every ~3.1
 * (measured: tests/tools/dfma_ops.cu), so per-lane coefficient registers would cap the filter at
 * ~70 % of the FP64 pipe. */
#define VS_FULL 0xffffffffu

/* ------------------------------------------------------------------------------------------------
 * conversions with x86-64 semantics (cvttsd2si: out-of-range -> 0x80000000, then keep low 16 bits)
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int32_t vs_d2i(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int32_t)0x80000000u;
    return __double2int_rz(v);
}
__device__ __forceinline__ int16_t vs_d2s(double v)
{
    return (int16_t)(uint16_t)(uint32_t)vs_d2i(v);
}
/* (short)ceil(v) for |v| < 2^31 (every pulse/noise value: |v| <= 1.8*32767 resp. NDW): one F2I */
__device__ __forceinline__ int vs_ceil_s16(double v)
{
    return (int)(int16_t)(uint16_t)(uint32_t)__double2int_ru(v);
}

/* r / d for a CONSTANT d with the reciprocal-multiply + FMA-residual sequence.  Exact (== IEEE
 * division) for every r in [0, 2^31) and both constants used here: checked exhaustively by
 * tests/tools/divcheck.c. */
__device__ __forceinline__ double vs_div_const(double r, double d, double inv)
{
    const double q0 = __dmul_rn(r, inv);
    const double rem = __fma_rn(-q0, d, r);
    return __fma_rn(rem, inv, q0);
}
#define VS_INV_RM  (1.0 / 2147483647.0)
#define VS_RM4     (2147483647.0 * 10000.0)
#define VS_INV_RM4 (1.0 / (2147483647.0 * 10000.0))

/* ------------------------------------------------------------------------------------------------
 * glibc random() TYPE_3 (r[i] = r[i-3] + r[i-31], output >> 1).  State lives in shared memory,
 * word-major ([31][VS_NT]) so that lanes never collide on a bank whatever their private index is.
 * ---------------------------------------------------------------------------------------------- */
struct VsRng {
    uint32_t *r;   /* shared memory base + threadIdx.x */
    int f;         /* front index; back index is f-3 (mod 31) */
};

template <int STRIDE = VS_NT>
__device__ __forceinline__ int32_t vs_rng_next(VsRng &g)
{
    const int b = g.f >= 3 ? g.f - 3 : g.f + 28;
    const uint32_t v = g.r[g.f * STRIDE] + g.r[b * STRIDE];
    g.r[g.f * STRIDE] = v;
    g.f = (g.f == VS_RNG_DEG - 1) ? 0 : g.f + 1;
    return (int32_t)(v >> 1);
}

template <int STRIDE = VS_NT>
__device__ void vs_rng_seed(VsRng &g, uint32_t seed)
{
    int32_t w = (int32_t)(seed ? seed : 1u);
    g.r[0] = (uint32_t)w;
    for (int i = 1; i < VS_RNG_DEG; i++) {          /* 16807*w mod (2^31-1), Schrage, signed */
        const int32_t hi = w / 127773, lo = w % 127773;
        w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        g.r[i * STRIDE] = (uint32_t)w;
    }
    g.f = 3;
    for (int i = 0; i < 310; i++) (void)vs_rng_next<STRIDE>(g);
}

/* store the state rotated so that a reader may assume f = 3: words 0..2 are the newest, word 3 the oldest */
template <int STRIDE = VS_NT>
__device__ void vs_rng_save(const VsRng &g, uint32_t *dst)
{
    int j = g.f >= 3 ? g.f - 3 : g.f + 28;
    for (int k = 0; k < VS_RNG_DEG; k++) {
        dst[k] = g.r[j * STRIDE];
        j = (j == VS_RNG_DEG - 1) ? 0 : j + 1;
    }
}
/* ------------------------------------------------------------------------------------------------
 * pulse samples (flowgen_shimmer.c:319, :328) and the noise sample (:387, :394, :591-600)
 *   rising : ceil((A*0.5)*(1-c)) == ceil(A*h) with h = 0.5*(1-c) tabulated (scaling by 0.5 is exact)
 *   falling: ceil(A*((K*c - K) + 1.0))
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int vs_rising(double Ad, double h)
{
    return vs_ceil_s16(__dmul_rn(Ad, h));
}
__device__ __forceinline__ int vs_falling(double Ad, double Kd, double c)
{
    return vs_ceil_s16(__dmul_rn(Ad, __dadd_rn(__dsub_rn(__dmul_rn(Kd, c), Kd), 1.0)));
}
__device__ __forceinline__ int vs_noise_w(int32_t r, int32_t ndw)
{
    const double u = vs_div_const((double)r, VS_RAND_MAX_D, VS_INV_RM);
    const double nd = (double)ndw;
    return vs_ceil_s16(__dsub_rn(__dmul_rn(u, nd), __dmul_rn(nd, 0.5)));
}
/* truncate((float)x + w): both are 16-bit integers, so the float sum is exact and ceil is a no-op */
__device__ __forceinline__ int vs_add_clip(int x, int w)
{
    const int s = x + w;
    return s > 32767 ? 32767 : (s < -32767 ? -32767 : s);
}

/* ================================================================================================
 * PLAN: one thread per stream
 * ============================================================================================== */
template <bool LOG>
__global__ void __launch_bounds__(VS_PLAN_NT, 1) vs_plan_kernel(const VsPlanArgs a)
{
    extern __shared__ __align__(16) uint32_t s_rng[];                  /* [31][VS_PLAN_NT] RNG states (+ padding up to VS_PLAN_SMEM) */
    const uint32_t s = blockIdx.x * VS_PLAN_NT + threadIdx.x;
    if (s >= a.n_streams) return;
    const VsStream st = a.streams[s];
    VsRng g;
    g.r = s_rng + threadIdx.x;
    vs_rng_seed<VS_PLAN_NT>(g, st.seed);                                             /* flowgen_shimmer.c:241 */

    const bool do_jit = (st.flags & VS_F_JITTER) && st.jitter != 0.0f;    /* :248 */
    const bool do_shm = (st.flags & VS_F_SHIMMER) && st.shimmer != 0.0f;  /* :295 */
    const bool noise = (st.flags & VS_F_NOISE) != 0;                      /* :373 */
    const bool pulse = LOG || noise;
    const int P = st.P, T2 = st.T2;
    const float Pf = (float)P, ampf = (float)st.amp;
    const float t_hi = __fmul_rn(1.2f, Pf), t_lo = __fmul_rn(0.8f, Pf);
    const float a_hi = __fmul_rn(1.8f, ampf), a_lo = __fmul_rn(0.2f, ampf);
    const double jit = (double)st.jitter, shm = (double)st.shimmer;
    const double jit2 = __dmul_rn(2.0, jit), shm2 = __dmul_rn(2.0, shm);
    const double P2 = __dmul_rn(2.0, (double)P), amp2 = __dmul_rn(2.0, (double)st.amp);
    const double Kbase = (double)st.K, kv2 = (double)__fmul_rn(2.0f, st.Kvar);
    const double *ht = a.costab + st.cos_off;      /* h[0..T2) = 0.5*(1-c), then c[0..T2) */
    const double *ct = ht + T2;
    const int DCi = (int)ceilf(st.DC);             /* (float)x < DC  <=>  x < ceil(DC) for integer x */
    const int DCs = st.DCs;

    int T = P, T4 = 0, ndw = 0;
    float dper = 0.0f, dsh = 0.0f;
    uint32_t count = 0, np = 0, next_c = 0;
    VsPeriod *tab = a.table + st.tab_off;
    VsChunk *chunks = a.chunks + st.chunk0;
    uint32_t next_target = st.n_chunks ? chunks[0].gen_target : 0xffffffffu;
    vs_period_rec *log = LOG ? (vs_period_rec *)a.log + st.log_off : nullptr;
    int guard = 0;

    do {
        uint32_t nd = 0;
        if (do_jit) {                                                     /* :276-290 */
            const double prev = (double)dper;
            float cur;
            do {
                const int32_t r = vs_rng_next<VS_PLAN_NT>(g); nd++;
                double t = vs_div_const((double)r, VS_RM4, VS_INV_RM4);
                t = __dmul_rn(__dmul_rn(t, 40000.0), jit);
                const double J = (double)__double2float_rn(__dsub_rn(t, jit2));
                const double den = __dsub_rn(2.0, J);
                const double q1 = __ddiv_rn(__dmul_rn(prev, __dadd_rn(2.0, J)), den);
                const double q2 = __ddiv_rn(__dmul_rn(P2, J), den);
                cur = __double2float_rn(__dadd_rn(q1, q2));
                T = (int)vs_d2s(ceil((double)__fadd_rn(Pf, cur)));
                if (++guard > (1 << 22)) { atomicExch(a.status, VS_ERANGE); return; }
            } while ((float)T > t_hi || (float)T < t_lo);
            dper = cur;
        }
        float A = ampf, S = 0.0f;
        if (do_shm) {                                                     /* :296-306 */
            const double prev = (double)dsh;
            float cur;
            do {
                const int32_t r = vs_rng_next<VS_PLAN_NT>(g); nd++;
                const float eps = __fmul_rn((float)r, 4.656612873077393e-10f);   /* / (float)RAND_MAX == * 2^-31, exact */
                S = __double2float_rn(__dsub_rn(__dmul_rn(__dmul_rn((double)eps, 4.0), shm), shm2));
                const double den = __dsub_rn(2.0, (double)S);
                const double q1 = __ddiv_rn(__dmul_rn(prev, __dadd_rn(2.0, (double)S)), den);
                const double q2 = __ddiv_rn(__dmul_rn(amp2, (double)S), den);
                cur = __double2float_rn(__dadd_rn(q1, q2));
                A = __fadd_rn(ampf, cur);
                if (++guard > (1 << 22)) { atomicExch(a.status, VS_ERANGE); return; }
            } while (A > a_hi || A < a_lo);
            dsh = cur;
        }
        if (T < 1 || T > 32767) { atomicExch(a.status, VS_ERANGE); return; }

        /* closure-speed draw, always consumed (:325) */
        const int32_t rk = vs_rng_next<VS_PLAN_NT>(g); nd++;
        const double kq = __dsub_rn(vs_div_const((double)rk, VS_RAND_MAX_D, VS_INV_RM), 0.5);
        const float Knew = __double2float_rn(__dmul_rn(Kbase, __dadd_rn(1.0, __dmul_rn(kv2, kq))));

        /* chunks whose generation starts inside this period: remember where we are */
        while (next_target < count + (uint32_t)T) {
            chunks[next_c].first_period = np;
            if (noise && a.rng_snap) vs_rng_save<VS_PLAN_NT>(g, a.rng_snap + (size_t)(st.chunk0 + next_c) * 32);
            next_c++;
            next_target = next_c < st.n_chunks ? chunks[next_c].gen_target : 0xffffffffu;
        }

        int T3 = 2 * T2;
        float x_pow = 0.0f, w_pow = 0.0f;
        uint32_t n_noise = 0;
        if (pulse) {
            /* one pass over the open phase: T4 = last rising index below DC (:320-323), T3 = first
             * falling index below DC (:329), and the float power sum over [T4,T3) in index order
             * (:374-378).  The sum restarts whenever T4 moves; if T4 never moves in this period the
             * sum that started at the stale T4 is the one the reference computes. */
            const double Ad = (double)A, Kd = (double)Knew;
            float aux_new = 0.0f, aux_old = 0.0f;
            bool moved = false;
            const int T4_old = T4;
            for (int i = 0; i < T2; i++) {
                int x = vs_rising(Ad, __ldg(ht + i));
                if (x < DCi) { x = DCs; T4 = i; moved = true; aux_new = 0.0f; }
                const float sq = __fmul_rn((float)x, (float)x);
                if (moved) aux_new = __fadd_rn(aux_new, sq);
                else if (i >= T4_old) aux_old = __fadd_rn(aux_old, sq);
            }
            float aux = moved ? aux_new : aux_old;
            int i;
            for (i = T2; i < 2 * T2; i++) {
                const int x = vs_falling(Ad, Kd, __ldg(ct + i - T2));
                if (x < DCi) break;
                aux = __fadd_rn(aux, __fmul_rn((float)x, (float)x));
            }
            T3 = i;
            if (noise) {                                                  /* :378-382 */
                const float span = __fsub_rn((float)T3, (float)T4);
                x_pow = __fdiv_rn(aux, span);
                const float ax = __double2float_rn(__dadd_rn(1.0, (double)__fdiv_rn(span, (float)T)));
                ndw = vs_d2i(sqrt((double)__fdiv_rn(__fmul_rn(__fmul_rn(12.0f, ax), x_pow), st.noise)));
                n_noise = (uint32_t)(T4 + (T > T3 ? T - T3 : 0));
                if (LOG) {
                    float wa = 0.0f;
                    for (uint32_t k = 0; k < n_noise; k++) {
                        const int w = vs_noise_w(vs_rng_next<VS_PLAN_NT>(g), ndw);
                        wa = __fadd_rn(wa, __fmul_rn((float)w, (float)w));
                    }
                    w_pow = __fdiv_rn(wa, (float)T);
                } else {
                    for (uint32_t k = 0; k < n_noise; k++) (void)vs_rng_next<VS_PLAN_NT>(g);
                }
            }
        }

        if (np >= st.tab_cap || nd > 65535u || T3 > 65535 || T4 > 65535) {
            atomicExch(a.status, np >= st.tab_cap ? VS_ENOMEM : VS_ERANGE);
            return;
        }
        VsPeriod e;
        e.Ad = (double)A; e.Kd = (double)Knew; e.start = count;
        e.T_np = (uint32_t)T | (nd << 16);
        e.T34 = (uint32_t)T3 | ((uint32_t)T4 << 16);
        e.ndw = ndw;
        tab[np] = e;
        if (LOG) {
            vs_period_rec r;
            r.T = T; r.T2 = T2; r.T3 = T3; r.T4 = T4; r.A = A; r.Knew = Knew; r.S = S;
            r.ndraws = (int32_t)(nd + n_noise); r.ndw = ndw; r.x_pow = x_pow; r.w_pow = w_pow; r.reserved = 0;
            r.start = count;
            log[np] = r;
        }
        count += (uint32_t)T;                                             /* :413 */
        np++;
    } while (count < st.n);                                               /* :423 */
    a.n_periods[s] = np;
}

/* ---- noise: the warp steps ONE row's random() 31 values at a time ------------------------------------
 * glibc TYPE_3 is r[n] = r[n-31] + r[n-3] (mod 2^32), output r[n] >> 1.  With lane l < 31 holding
 * st[l] = r[n-31+l], the next 31 words are prefix sums along the three stride-3 chains:
 *     new[l] = st[l] + (l < 3 ? st[28+l] : new[l-3])
 * i.e. an inclusive scan with shuffle distances 3, 6, 12, 24 (chains are at most 11 long). */
__device__ __forceinline__ uint32_t vs_rng_round(uint32_t st, int lane)
{
    const uint32_t wrap = __shfl_sync(VS_FULL, st, (lane + 28) & 31);
    uint32_t v = st + (lane < 3 ? wrap : 0u), u;
    u = __shfl_up_sync(VS_FULL, v, 3);  if (lane >= 3)  v += u;
    u = __shfl_up_sync(VS_FULL, v, 6);  if (lane >= 6)  v += u;
    u = __shfl_up_sync(VS_FULL, v, 12); if (lane >= 12) v += u;
    u = __shfl_up_sync(VS_FULL, v, 24); if (lane >= 24) v += u;
    return v;
}

/* srandom() for the warp-cooperative generator: every lane runs the 30 Schrage steps and keeps the word that
 * is (lane)-th oldest once f = 3, i.e. word (lane+3) mod 31; the 310 discarded values are 10 rounds. */
__device__ __forceinline__ uint32_t vs_rng_seed_warp(uint32_t seed, int lane)
{
    int32_t w = (int32_t)(seed ? seed : 1u);
    const int want = (lane + 3) % VS_RNG_DEG;
    uint32_t rs = (uint32_t)w;
    for (int i = 1; i < VS_RNG_DEG; i++) {
        const int32_t hi = w / 127773, lo = w % 127773;
        w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        if (i == want) rs = (uint32_t)w;
    }
    for (int k = 0; k < 10; k++) rs = vs_rng_round(rs, lane);
    return rs;
}

/* advance the row's generator by m values; they go to out[0..m) when out is not NULL.  Returns the new
 * state (lanes 0..30).  m is warp-uniform. */
__device__ __forceinline__ uint32_t vs_rng_gen(uint32_t st, int m, int lane, int32_t *out)
{
    int done = 0;
    for (; m - done >= VS_RNG_DEG; done += VS_RNG_DEG) {
        st = vs_rng_round(st, lane);
        if (out && lane < VS_RNG_DEG) out[done + lane] = (int32_t)(st >> 1);
    }
    const int u = m - done;
    if (u > 0) {                                     /* part of a round: the state window slides by u words */
        const uint32_t nx = vs_rng_round(st, lane);
        if (out && lane < u) out[done + lane] = (int32_t)(nx >> 1);
        const int src = lane + u;
        const uint32_t keep = __shfl_sync(VS_FULL, st, src & 31);
        const uint32_t fresh = __shfl_sync(VS_FULL, nx, (src - VS_RNG_DEG) & 31);
        st = src < VS_RNG_DEG ? keep : fresh;
    }
    return st;
}

/* ================================================================================================
 * PLAN, one WARP per stream: for small batches (a single ten-minute stream is 71 k sequential periods)
 * and for streams with glottal noise, where one thread per stream leaves the GPU idle or crawls through
 * hundreds of serial pulse samples and noise draws per period.
 *   - random(): 31 values per round by the whole warp (vs_rng_round);
 *   - every value's three possible meanings (jitter draw, shimmer draw, closure-speed draw; they do not
 *     depend on the walk's state) are worked out lane-parallel, one value per lane, incl. the two
 *     divisions whose numerator is state-free and a refined reciprocal for the one that is not;
 *   - the walk itself (flowgen_shimmer.c:276-325) runs warp-uniform, reading those values back by
 *     broadcast; its remaining division is the reciprocal-multiply sequence with an exactness check and
 *     an IEEE division when the check is inconclusive;
 *   - with noise: the open phase is evaluated 32 samples per step, the float power sum then runs in
 *     index order over the 32 squares (the sum is not associative), and the period's noise draws are
 *     stepped over with vs_rng_gen.
 * Same tables, chunk marks and RNG snapshots as vs_plan_kernel<false>.
 * ============================================================================================== */
#define VS_PLANW_NT 128

/* n/d given y ~ 1/d; sets doubt unless the result is provably the correctly rounded quotient */
__device__ __forceinline__ double vs_div_checked_y(double n, double d, double y, bool &doubt)
{
    const double q0 = __dmul_rn(n, y);
    const double q = __fma_rn(__fma_rn(-q0, d, n), y, q0);
    const double rem = __fma_rn(-q, d, n);
    const int e = (__double2hiint(q) >> 20) & 0x7ff;                /* ulp(q)/2 = 2^(e-53-1023) */
    const double half_ulp = __hiloint2double((e - 53) << 20, 0);
    const bool sure = (e > 60 && e < 0x7fe) && (fabs(rem) < __dmul_rn(half_ulp, fabs(d)));
    doubt = !sure && !(q == 0.0 && rem == 0.0);
    return q;
}
__device__ __forceinline__ double vs_recip_refined(double d)
{
    double y = (double)__frcp_rn((float)d);
    y = __fma_rn(y, __fma_rn(-d, y, 1.0), y);
    y = __fma_rn(y, __fma_rn(-d, y, 1.0), y);
    y = __fma_rn(y, __fma_rn(-d, y, 1.0), y);
    return y;
}

__global__ void __launch_bounds__(VS_PLANW_NT) vs_plan_warp_kernel(const VsPlanArgs a)
{
    /* per random() value, in a ring of 64 slots (two rounds of 31 are outstanding at most):
     * 2+J, 2-J, 1/(2-J), 2P*J/(2-J); the same four for S; Knew */
    __shared__ double s_itp[VS_PLANW_NT / 32][8][64];
    __shared__ float s_kn[VS_PLANW_NT / 32][64];
    __shared__ __align__(16) float s_sq[VS_PLANW_NT / 32][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t s = blockIdx.x * (VS_PLANW_NT / 32) + wib;
    if (s >= a.n_streams) return;
    const VsStream st = a.streams[s];
    double (*itp)[64] = s_itp[wib];
    float *kn = s_kn[wib];
    float *sqb = s_sq[wib];

    const bool do_jit = (st.flags & VS_F_JITTER) && st.jitter != 0.0f;    /* :248 */
    const bool do_shm = (st.flags & VS_F_SHIMMER) && st.shimmer != 0.0f;  /* :295 */
    const bool noise = (st.flags & VS_F_NOISE) != 0;                      /* :373 */
    const int P = st.P, T2 = st.T2;
    const float Pf = (float)P, ampf = (float)st.amp;
    const float t_hi = __fmul_rn(1.2f, Pf), t_lo = __fmul_rn(0.8f, Pf);
    const float a_hi = __fmul_rn(1.8f, ampf), a_lo = __fmul_rn(0.2f, ampf);
    const double jit = (double)st.jitter, shm = (double)st.shimmer;
    const double jit2 = __dmul_rn(2.0, jit), shm2 = __dmul_rn(2.0, shm);
    const double P2 = __dmul_rn(2.0, (double)P), amp2 = __dmul_rn(2.0, (double)st.amp);
    const double Kbase = (double)st.K, kv2 = (double)__fmul_rn(2.0f, st.Kvar);
    const double *ht = a.costab + st.cos_off;
    const double *ct = ht + T2;
    const int DCi = (int)ceilf(st.DC);
    const int DCs = st.DCs;

    uint32_t rs = vs_rng_seed_warp(st.seed, lane);                       /* flowgen_shimmer.c:241 */

    /* Value source.  Values are numbered from the last reset; value i lives in ring slot i & 63.
     * rs = generator state before round A; nxa = round A = values [base, base+31); nxb = round B =
     * values [base+31, base+62) once have_b; pos = next value to use. */
    uint32_t nxa = 0, nxb = 0;
    int base = 0, pos = 0;
    bool have_b = false;
    auto interpret = [&](uint32_t words, int first) {     /* lane-parallel: what each value means as J, S, K draw */
        __syncwarp();
        if (lane < VS_RNG_DEG) {
            const int32_t r = (int32_t)(words >> 1);
            const int slot = (first + lane) & 63;
            if (do_jit) {                                                  /* :277-281 */
                double t = vs_div_const((double)r, VS_RM4, VS_INV_RM4);
                t = __dmul_rn(__dmul_rn(t, 40000.0), jit);
                const double J = (double)__double2float_rn(__dsub_rn(t, jit2));
                const double den = __dsub_rn(2.0, J), y = vs_recip_refined(den), num = __dmul_rn(P2, J);
                bool doubt;
                double q = vs_div_checked_y(num, den, y, doubt);
                if (doubt) q = __ddiv_rn(num, den);
                itp[0][slot] = __dadd_rn(2.0, J);
                itp[1][slot] = den;
                itp[2][slot] = y;
                itp[3][slot] = q;
            }
            if (do_shm) {                                                  /* :297-301 */
                const float eps = __fmul_rn((float)r, 4.656612873077393e-10f);
                const double S = (double)__double2float_rn(__dsub_rn(__dmul_rn(__dmul_rn((double)eps, 4.0), shm), shm2));
                const double den = __dsub_rn(2.0, S), y = vs_recip_refined(den), num = __dmul_rn(amp2, S);
                bool doubt;
                double q = vs_div_checked_y(num, den, y, doubt);
                if (doubt) q = __ddiv_rn(num, den);
                itp[4][slot] = __dadd_rn(2.0, S);
                itp[5][slot] = den;
                itp[6][slot] = y;
                itp[7][slot] = q;
            }
            const double kq = __dsub_rn(vs_div_const((double)r, VS_RAND_MAX_D, VS_INV_RM), 0.5);   /* :325 */
            kn[slot] = __double2float_rn(__dmul_rn(Kbase, __dadd_rn(1.0, __dmul_rn(kv2, kq))));
        }
        __syncwarp();
    };
    auto restart = [&](uint32_t state) {                  /* value numbering starts again at this generator state */
        rs = state;
        nxa = vs_rng_round(rs, lane);
        base = 0; pos = 0; have_b = false;
        interpret(nxa, 0);
    };
    auto make_b = [&]() {
        nxb = vs_rng_round(nxa, lane);
        interpret(nxb, base + VS_RNG_DEG);
        have_b = true;
    };
    auto refill = [&](bool ahead) {                       /* afterwards pos < base+31, and round B exists if `ahead` */
        while (pos >= base + VS_RNG_DEG) {
            if (!have_b) make_b();
            rs = nxa; nxa = nxb; base += VS_RNG_DEG; have_b = false;
        }
        if (ahead && !have_b) make_b();
    };
    auto slide = [&](uint32_t before, uint32_t words, int u) -> uint32_t {   /* state after u (0..31) values of a round */
        const int src = lane + u;
        const uint32_t keep = __shfl_sync(VS_FULL, before, src & 31);
        const uint32_t fr = __shfl_sync(VS_FULL, words, (src - VS_RNG_DEG) & 31);
        return src < VS_RNG_DEG ? keep : fr;
    };
    auto state_now = [&]() -> uint32_t {                  /* lanes 0..30, oldest word first */
        const int off = pos - base;                       /* 0..62 */
        return off <= VS_RNG_DEG ? slide(rs, nxa, off) : slide(nxa, nxb, off - VS_RNG_DEG);
    };
    restart(rs);

    int T = P, T4 = 0, ndw = 0;
    float dper = 0.0f, dsh = 0.0f;
    uint32_t count = 0, np = 0, next_c = 0;
    VsPeriod *tab = a.table + st.tab_off;
    VsChunk *chunks = a.chunks + st.chunk0;
    uint32_t next_target = st.n_chunks ? chunks[0].gen_target : 0xffffffffu;
    /* the target after that is fetched one mark ahead: a single long stream crosses a chunk start every few
     * periods and would otherwise wait for a dependent global load each time */
    uint32_t after_target = st.n_chunks > 1 ? chunks[1].gen_target : 0xffffffffu;
    int guard = 0;
    const bool both = do_jit && do_shm;

    do {
        /* Tight loop for the common stream (jitter and shimmer on, no noise) while everything goes the common
         * way: three values per period straight off the ring, both draws accepted, divisions conclusive.  It
         * leaves to the general period below for anything else -- values running out (refill), a rejected
         * draw, an inconclusive division -- and comes back afterwards. */
        if (both && !noise) {
            const int lim = base + (have_b ? 2 * VS_RNG_DEG : VS_RNG_DEG);
            while (pos + 3 <= lim && count < st.n && np < st.tab_cap) {
                const int pj = pos & 63, ps = (pos + 1) & 63, pk = (pos + 2) & 63;
                const double numj = __dmul_rn((double)dper, itp[0][pj]);
                const double nums = __dmul_rn((double)dsh, itp[4][ps]);
                bool dj, ds;
                const double qj = vs_div_checked_y(numj, itp[1][pj], itp[2][pj], dj);
                const double qs = vs_div_checked_y(nums, itp[5][ps], itp[6][ps], ds);
                const float curJ = __double2float_rn(__dadd_rn(qj, itp[3][pj]));
                const float curS = __double2float_rn(__dadd_rn(qs, itp[7][ps]));
                const float Tf = ceilf(__fadd_rn(Pf, curJ));
                const float An = __fadd_rn(ampf, curS);
                if (dj || ds || Tf > t_hi || Tf < t_lo || !(Tf >= 1.0f && Tf <= 32767.0f) || An > a_hi || An < a_lo) break;
                dper = curJ; dsh = curS; T = (int)Tf;
                pos += 3;
                while (next_target < count + (uint32_t)T) {
                    if (lane == 0) chunks[next_c].first_period = np;
                    next_c++;
                    next_target = after_target;
                    after_target = next_c + 1 < st.n_chunks ? chunks[next_c + 1].gen_target : 0xffffffffu;
                }
                if (lane == 0) {
                    VsPeriod e;
                    e.Ad = (double)An; e.Kd = (double)kn[pk]; e.start = count;
                    e.T_np = (uint32_t)T | (3u << 16);
                    e.T34 = (uint32_t)(2 * T2);
                    e.ndw = 0;
                    tab[np] = e;
                }
                count += (uint32_t)T;
                np++;
            }
            if (count >= st.n) break;
        }
        uint32_t nd = 0;
        float A = ampf, Knew = 0.0f;
        bool committed = false;
        /* without noise the values of consecutive periods are consecutive: keep a round in reserve so that
         * the common case never meets a round boundary */
        if (pos >= base + VS_RNG_DEG || !(noise || have_b)) refill(!noise);
        /* Common case first: jitter and shimmer both on, every draw accepted at once.  The two random walks
         * are independent recurrences, so their division chains run side by side; acceptance is tested on
         * ceilf(P + dPer), the same integer as (short)ceil((double)..) when it is in range.  Anything else
         * -- a rejected draw, an inconclusive division check -- takes the literal loops below from pos. */
        if (both && (have_b || pos + 3 <= base + VS_RNG_DEG)) {
            const int pj = pos & 63, ps = (pos + 1) & 63, pk = (pos + 2) & 63;
            const double numj = __dmul_rn((double)dper, itp[0][pj]);
            const double nums = __dmul_rn((double)dsh, itp[4][ps]);
            bool dj, ds;
            const double qj = vs_div_checked_y(numj, itp[1][pj], itp[2][pj], dj);
            const double qs = vs_div_checked_y(nums, itp[5][ps], itp[6][ps], ds);
            const float curJ = __double2float_rn(__dadd_rn(qj, itp[3][pj]));
            const float curS = __double2float_rn(__dadd_rn(qs, itp[7][ps]));
            const float Tf = ceilf(__fadd_rn(Pf, curJ));
            const float An = __fadd_rn(ampf, curS);
            const bool ok = !dj && !ds && !(Tf > t_hi || Tf < t_lo) && Tf >= 1.0f && Tf <= 32767.0f && !(An > a_hi || An < a_lo);
            if (ok) {
                dper = curJ; dsh = curS; T = (int)Tf; A = An;
                Knew = kn[pk];
                nd = 3u;
                pos += 3;
                committed = true;
            }
        }
        if (!committed) {
            if (do_jit) {                                                     /* :276-290 */
                const double prev = (double)dper;
                float cur;
                for (;;) {
                    refill(false);
                    const int p = pos++ & 63; nd++;
                    const double num = __dmul_rn(prev, itp[0][p]), den = itp[1][p];
                    bool doubt;
                    double q1 = vs_div_checked_y(num, den, itp[2][p], doubt);
                    if (doubt) q1 = __ddiv_rn(num, den);
                    cur = __double2float_rn(__dadd_rn(q1, itp[3][p]));
                    T = (int)vs_d2s(ceil((double)__fadd_rn(Pf, cur)));
                    if (++guard > (1 << 22)) { atomicExch(a.status, VS_ERANGE); return; }
                    if (!((float)T > t_hi || (float)T < t_lo)) break;
                }
                dper = cur;
            }
            if (do_shm) {                                                     /* :296-306 */
                const double prev = (double)dsh;
                float cur;
                for (;;) {
                    refill(false);
                    const int p = pos++ & 63; nd++;
                    const double num = __dmul_rn(prev, itp[4][p]), den = itp[5][p];
                    bool doubt;
                    double q1 = vs_div_checked_y(num, den, itp[6][p], doubt);
                    if (doubt) q1 = __ddiv_rn(num, den);
                    cur = __double2float_rn(__dadd_rn(q1, itp[7][p]));
                    A = __fadd_rn(ampf, cur);
                    if (++guard > (1 << 22)) { atomicExch(a.status, VS_ERANGE); return; }
                    if (!(A > a_hi || A < a_lo)) break;
                }
                dsh = cur;
            }
            if (T < 1 || T > 32767) { atomicExch(a.status, VS_ERANGE); return; }
            refill(false);                                                    /* closure-speed draw (:325) */
            Knew = kn[pos++ & 63]; nd++;
        }

        while (next_target < count + (uint32_t)T) {
            if (lane == 0) chunks[next_c].first_period = np;
            if (noise && a.rng_snap) {
                const uint32_t now = state_now();
                if (lane < VS_RNG_DEG) a.rng_snap[(size_t)(st.chunk0 + next_c) * 32 + (lane + 3) % VS_RNG_DEG] = now;
            }
            next_c++;
            next_target = after_target;
            after_target = next_c + 1 < st.n_chunks ? chunks[next_c + 1].gen_target : 0xffffffffu;
        }

        int T3 = 2 * T2;
        if (noise) {
            const double Ad = (double)A, Kd = (double)Knew;
            float aux = 0.0f;
            /* rising branch (:318-323): T4 = last index below DC; the power sum covers [T4, T2) whether T4
             * moved in this period or is the stale one */
            for (int b0 = 0; b0 < T2; b0 += 32) {
                const int i = b0 + lane;
                const bool valid = i < T2;
                int x = valid ? vs_rising(Ad, __ldg(ht + i)) : 0;
                const bool below = valid && x < DCi;
                const uint32_t m = __ballot_sync(VS_FULL, below);
                if (m) { T4 = b0 + 31 - __clz((int)m); aux = 0.0f; }
                if (below) x = DCs;
                sqb[lane] = (valid && i >= T4) ? __fmul_rn((float)x, (float)x) : 0.0f;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float4 v = reinterpret_cast<const float4 *>(sqb)[k];
                    aux = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(aux, v.x), v.y), v.z), v.w);
                }
                __syncwarp();
            }
            /* falling branch (:327-332): stops at the first index below DC = T3 */
            for (int b0 = T2; b0 < 2 * T2; b0 += 32) {
                const int i = b0 + lane;
                const bool valid = i < 2 * T2;
                const int x = valid ? vs_falling(Ad, Kd, __ldg(ct + i - T2)) : 0;
                const uint32_t m = __ballot_sync(VS_FULL, valid && x < DCi);
                const int first = m ? __ffs((int)m) - 1 : 32;
                sqb[lane] = (valid && lane < first) ? __fmul_rn((float)x, (float)x) : 0.0f;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float4 v = reinterpret_cast<const float4 *>(sqb)[k];
                    aux = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(aux, v.x), v.y), v.z), v.w);
                }
                __syncwarp();
                if (m) { T3 = b0 + first; break; }
            }
            /* :378-382 */
            const float span = __fsub_rn((float)T3, (float)T4);
            const float x_pow = __fdiv_rn(aux, span);
            const float ax = __double2float_rn(__dadd_rn(1.0, (double)__fdiv_rn(span, (float)T)));
            ndw = vs_d2i(sqrt((double)__fdiv_rn(__fmul_rn(__fmul_rn(12.0f, ax), x_pow), st.noise)));
            const uint32_t n_noise = (uint32_t)(T4 + (T > T3 ? T - T3 : 0));
            /* the period's noise values are the render kernel's business: step over them */
            restart(vs_rng_gen(state_now(), (int)n_noise, lane, nullptr));
        }

        if (np >= st.tab_cap || (!committed && nd > 65535u) || (noise && (T3 > 65535 || T4 > 65535))) {
            atomicExch(a.status, np >= st.tab_cap ? VS_ENOMEM : VS_ERANGE);
            return;
        }
        if (lane == 0) {
            VsPeriod e;
            e.Ad = (double)A; e.Kd = (double)Knew; e.start = count;
            e.T_np = (uint32_t)T | (nd << 16);
            e.T34 = (uint32_t)T3 | ((uint32_t)T4 << 16);
            e.ndw = ndw;
            tab[np] = e;
        }
        count += (uint32_t)T;                                             /* :413 */
        np++;
    } while (count < st.n);                                               /* :423 */
    if (lane == 0) a.n_periods[s] = np;
}

/* ================================================================================================
 * RENDER: one lane per (stream, chunk), one warp per 32 of them
 * ============================================================================================== */
enum { VS_MODE_FLOW = 0, VS_MODE_SYNTH = 1, VS_MODE_FILTER = 2 };

/* what the other lanes need to know about a lane's chunk; lives in shared memory */
struct __align__(16) VsLane {
    const VsPeriod *tab;
    const double *ct;        /* c[0..T2), h[0..T2) */
    int16_t *orow;
    const int16_t *fin;
    int32_t nstart, lo, hi, blk0;
    int32_t T2, DCi, DCs, noise;
    uint32_t tab_cap;        /* entries of the row's period table: indices are checked against it */
    uint32_t chunk;          /* chunk id (noise: where the row's RNG snapshot lives)              */
    uint32_t q0;             /* the chunk's first period: its perturbation draws are already consumed */
    uint32_t pad;
};

struct VsEnt {               /* a period-table entry in registers */
    double Ad, Kd;
    int start, T, T3, T4, ndw, npert;
};

__device__ __forceinline__ VsEnt vs_load_entry(const VsPeriod *p)
{
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const int4 b = __ldg(reinterpret_cast<const int4 *>(p) + 1);
    VsEnt e;
    e.Ad = a.x; e.Kd = a.y;
    e.start = b.x;
    e.T = b.y & 0xffff; e.npert = (int)((uint32_t)b.y >> 16);
    e.T3 = b.z & 0xffff; e.T4 = (int)((uint32_t)b.z >> 16);
    e.ndw = b.w;
    return e;
}

/* vowel_new.c:413-427, literally (exact mode) */
__device__ __forceinline__ int vs_round2int(double v)
{
    const double dec = __dsub_rn(v, floor(v));
    if (dec > 0.5) v = __dadd_rn(v, 1.0);
    if (v > 32767.0) v = 32767.0;
    else if (v < -32767.0) v = -32767.0;
    return (int)vs_d2s(floor(v));
}

/* Fast quantiser: round2int(v) == clip(ceil(v - 0.5)) (round half DOWN) except for v within one ulp
 * below an integer, where the reference's own x+1 rounds up.  One FP64 add in round-up mode with
 * the magic constant 2^51+2^50-0.5 (ulp there is 0.5) leaves 2*h in the low mantissa word, h the
 * smallest multiple of 0.5 >= v-0.5; ceil(h) = (2h+1)>>1.  The high word tells when |v| >= 2^30. */
__device__ __forceinline__ int vs_quant_fast(double v)
{
    const double s = __dadd_ru(v, 3377699720527871.5);
    const int k2 = __double2loint(s);
    const int hw = __double2hiint(s);
    int k = (k2 + 1) >> 1;
    k = max(-32767, min(32767, k));
    if (hw - 0x43280000 != (k2 >> 31)) k = __double2hiint(v) < 0 ? -32767 : 32767;
    return k;
}

/* same, for batches whose waveform is bounded below 2^30 (host-side bound on gain * l1 gain of the preset) */
__device__ __forceinline__ int vs_quant_fast_nocheck(double v)
{
    const int k2 = __double2loint(__dadd_ru(v, 3377699720527871.5));
    return max(-32767, min(32767, (k2 + 1) >> 1));
}

/* ---- shared-memory geometry of the render kernel ------------------------------------------------
 * A CTA serves VS_NP groups of 32 rows (one row = one stream-chunk).  In the filtering modes a group
 * is worked by one CONSUMER warp (F phase; warps 0..NP-1, one per SM sub-partition) and VS_PW
 * PRODUCER warps (G and W phases) over two tiles; in flow mode by the VS_PW producer warps alone.  */
#define VS_NP        4
#define VS_PW        4
#define VS_MAXSEG    4                     /* period segments a row can queue per bookkeeping pass */
/* Samples per row and window.  Every window costs a fixed amount on both sides (barrier, restart of the
 * consumer's software pipeline, a bookkeeping pass, per-row loops, open phases cut in two), measured at
 * ~22 % of the time with 192-sample windows; the fused kernels have the shared memory for more:
 * 240 (30 16-byte pieces per row: one write-out step per row) measured as good as 288.
 * Flow-only mode (two CTAs per SM) stays at 192. */
__host__ __device__ constexpr int vs_win(int mode, bool noise) { return mode != 0 /* VS_MODE_FLOW */ ? VS_WIN_WIDE : VS_WIN; }
/* tile row stride in int16.  With a consumer (lane = row walks down a column) an odd number of words keeps the
 * columns conflict-free; flow-only mode has no column access, so its rows are 16-byte aligned instead and the
 * fill and write-out phases move 16 bytes per lane */
#define VS_TS_OF(MODE, WIN) ((MODE) == 0 /* VS_MODE_FLOW */ ? (WIN) + 8 : (WIN) + 2)
#define VS_TILE_I16_OF(MODE, WIN) (32 * VS_TS_OF(MODE, WIN))
#define VS_THREADS_PAIRED ((VS_NP + VS_NP * VS_PW) * 32)
/* flow mode: no consumer; the VS_PW warps of a group each do G then W for their own rows (one warp per 32
 * rows left the SM with 4-8 warps: 0.40 ms on the bench workload against 0.21 ms this way) */
#define VS_THREADS_FLOW   (VS_NP * VS_PW * 32)
#define VS_RENDER_THREADS(MODE) ((MODE) != VS_MODE_FLOW ? VS_THREADS_PAIRED : VS_THREADS_FLOW)

/* one pitch period's share of one row's window: everything the cooperative evaluation needs (48 B).
 * The evaluation counts samples k = 0.. from the segment's first in-window sample a0. */
struct __align__(16) VsSeg {
    double Ad, Kd;
    const double *tab;   /* table entry of sample a0 (h[0..T2) then c[0..T2) of the row's T2) */
    int16_t *out;        /* tile address of sample a0 */
    uint32_t nr;         /* low 16 bits: open-phase samples to evaluate from a0 on; high 16: how many of them rise */
    int DCi;
    int a0, a1;          /* in-period indices [a0,a1) that fall into this window (noise pass) */
};
#define VS_ITEM_SAMPLES 64                 /* open-phase samples per work item: two per lane */

/* NTILE = tiles per group: 2 (double buffer) in the filtering modes, 1 in flow mode */
#define VS_SMEM_TILES(NTILE, WIN) (VS_NP * (NTILE) * VS_TILE_I16_OF((NTILE) == 1 ? 0 : 1, WIN) * 2)
#define VS_SMEM_LANES (VS_NP * 32 * (int)sizeof(VsLane))
#define VS_SMEM_SEGS  (VS_NP * 32 * VS_MAXSEG * (int)sizeof(VsSeg))
#define VS_SEGX       3                                                /* noise words per segment: T3|T4, NoiseDistWidth, draws to skip */
#define VS_SMEM_NSEG  (VS_NP * 32 * 4 + VS_NP * 32 * VS_MAXSEG * VS_SEGX * 4)   /* per-row segment counts + noise words */
#define VS_ITEMS_BYTES (((32 * VS_MAXSEG * 4 + 4) * 2 + 15) & ~15)      /* work-item list of one producer warp */
#define VS_SMEM_ITEMS(NPROD) (VS_NP * (NPROD) * VS_ITEMS_BYTES)
#define VS_SMEM_BASE(NTILE, WIN) (VS_SMEM_TILES(NTILE, WIN) + VS_SMEM_LANES + VS_SMEM_SEGS + VS_SMEM_NSEG + VS_SMEM_ITEMS(VS_PW))
/* noise: one RNG state per row, [NP][32 rows][32 words], oldest word first, and a
 * scratch of random() values per producer warp */
#define VS_DRAW_SCRATCH 288                /* random() values of one period's share of a window (<= VS_WIN_WIDE) plus a few to skip */
#define VS_SMEM_NOISE (VS_NP * 32 * 32 * 4 + VS_NP * VS_PW * VS_DRAW_SCRATCH * 4)

__device__ __forceinline__ void vs_named_barrier(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

/* ---- G: generate window w of the rows {row0, row0+step, ...} of a tile --------------------------------
 * Given the period table a sample depends only on its index.  Closed phase = DC: a word fill.  The
 * open phase of a period (rising then falling branch, flowgen_shimmer.c:318-332) is one contiguous
 * index run.  Bookkeeping is lane-parallel (lane l <-> row row0 + l*step): each lane walks its row's
 * period table, queues segment descriptors in shared memory and expands them into a flat list of
 * work items (one item = 32 consecutive open-phase samples of one segment).  The warp then evaluates
 * four items per step, straight-line, so that four table loads and FP64 chains are in flight. */
#define VS_ITEM(row, seg, grp) ((uint16_t)(((row) * VS_MAXSEG + (seg)) | ((grp) << 7)))

template <int MODE, bool NOISE>
__device__ __forceinline__ void vs_gen_tile(int16_t *tile, const VsLane *lanes, VsSeg *segs, int *nsegs, uint16_t *items,
                                            int w, int lane, int row0, int step, const VsLane &mine, uint32_t &q,
                                            uint32_t *rngrow, int32_t *scratch, bool general)
{
    constexpr int WIN = vs_win(MODE, NOISE), TS = VS_TS_OF(MODE, WIN);
    const int myrow = row0 + lane * step;
    const bool have_row = myrow < 32;
    uint32_t *segx = reinterpret_cast<uint32_t *>(nsegs + 32);     /* [32][VS_MAXSEG][VS_SEGX] noise words per queued segment */

    if (MODE == VS_MODE_FILTER) {
        for (int j = row0; j < 32; j += step) {
            const VsLane L = lanes[j];
            const int wb = L.blk0 + w * WIN;
            if (wb >= L.hi) continue;
            int16_t *trow = tile + j * TS;
#pragma unroll
            for (int k = lane; k < WIN; k += 32) {
                const int m = wb + k;
                trow[k] = (m >= L.nstart && m < L.hi) ? __ldg(L.fin + m) : (int16_t)0;
            }
        }
        return;
    }

    /* 1. closed phase everywhere (:334-336), zeros outside the stream */
    for (int j = row0; j < 32; j += step) {
        const int wb = lanes[j].blk0 + w * WIN, hi = lanes[j].hi, ns = lanes[j].nstart;
        if (wb >= hi) continue;
        int16_t *trow = tile + j * TS;
        const int dcs = lanes[j].DCs;
        if (ns <= wb && wb + WIN <= hi) {
            const uint32_t pat = (uint32_t)(uint16_t)dcs * 0x10001u;
            if (MODE == VS_MODE_FLOW) {                                 /* 16-byte aligned rows */
                uint4 *t128 = reinterpret_cast<uint4 *>(trow);
#pragma unroll
                for (int k = lane; k < WIN / 8; k += 32) t128[k] = make_uint4(pat, pat, pat, pat);
            } else {
                uint32_t *t32 = reinterpret_cast<uint32_t *>(trow);
#pragma unroll
                for (int k = lane; k < WIN / 2; k += 32) t32[k] = pat;
            }
        } else {
#pragma unroll
            for (int k = lane; k < WIN; k += 32) {
                const int m = wb + k;
                trow[k] = (m >= ns && m < hi) ? (int16_t)dcs : (int16_t)0;
            }
        }
    }

    /* 2. bookkeeping passes + cooperative evaluation */
    const int wb_m = mine.blk0 + w * WIN;
    const int glo = max(wb_m, mine.nstart), ghi = min(wb_m + WIN, mine.hi);
    bool pending = have_row && wb_m < mine.hi;
    while (__any_sync(VS_FULL, pending)) {
        int n = 0, ngrp = 0;
        uint32_t grp_counts = 0;                              /* 4 bits per segment */
        if (pending) {
            VsSeg *my = segs + myrow * VS_MAXSEG;
            const int open_end = 2 * mine.T2;
#pragma unroll 1
            for (int sidx = 0; sidx < VS_MAXSEG; sidx++) {
                if (q >= mine.tab_cap) { pending = false; break; }          /* never walk past the row's table */
                const VsEnt e = vs_load_entry(mine.tab + q);
                const int pend = e.start + e.T;
                if (pend > glo) {
                    const int a0 = max(0, glo - e.start), a1 = min(e.T, ghi - e.start);
                    /* A above 32767: x[T2] = (short)ceil(A) is negative, the reference leaves the falling
                     * branch at once (:329) */
                    const int nopen = min(a1, e.Ad > 32767.0 ? mine.T2 : open_end) - a0;
                    if (nopen > 0 || (NOISE && mine.noise)) {
                        VsSeg sg;
                        sg.Ad = e.Ad; sg.Kd = e.Kd; sg.tab = mine.ct + a0;
                        sg.out = tile + myrow * TS + (e.start - wb_m) + a0;
                        sg.nr = (uint32_t)max(nopen, 0) | ((uint32_t)min(max(mine.T2 - a0, 0), WIN) << 16);
                        sg.DCi = mine.DCi;
                        sg.a0 = a0; sg.a1 = a1;
                        my[n] = sg;
                        if (NOISE) {
                            uint32_t *x = segx + (myrow * VS_MAXSEG + n) * VS_SEGX;
                            x[0] = (uint32_t)e.T3 | ((uint32_t)e.T4 << 16);
                            x[1] = (uint32_t)e.ndw;
                            /* a period that begins in this window draws its perturbations and K before
                             * its first noise sample (:283,:298,:325); the chunk's first period has
                             * them behind it already (the snapshot is taken after its K draw) */
                            x[2] = (a0 == 0 && q != mine.q0) ? (uint32_t)e.npert : 0u;
                        }
                        const int gcnt = nopen > 0 ? (nopen + VS_ITEM_SAMPLES - 1) / VS_ITEM_SAMPLES : 0;
                        grp_counts |= (uint32_t)gcnt << (4 * n);
                        ngrp += gcnt;
                        n++;
                    }
                }
                if (pend >= ghi) {                       /* window covered: q -> period of the next window's first sample */
                    if (pend == ghi) q++;
                    pending = false;
                    break;
                }
                q++;
            }
        }
        if (NOISE && have_row) nsegs[myrow] = n;
        /* flat work list: exclusive prefix sum of the per-row group counts */
        int incl = ngrp;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(VS_FULL, incl, o);
            if (lane >= o) incl += up;
        }
        const int total = __shfl_sync(VS_FULL, incl, 31);
        {
            int pos = incl - ngrp;
            uint16_t last = 0;
            for (int sidx = 0; sidx < n; sidx++) {
                const int gcnt = (int)((grp_counts >> (4 * sidx)) & 15u);
                for (int gi = 0; gi < gcnt; gi++) { last = VS_ITEM(myrow, sidx, gi); items[pos++] = last; }
            }
            /* pad to a multiple of 4 with copies of the last item: re-evaluating it is idempotent, so the
             * evaluation loop needs no validity test */
            if (ngrp > 0 && incl == total) { items[total] = last; items[total + 1] = last; items[total + 2] = last; }
        }
        __syncwarp();

        /* The table values of the NEXT four items are requested before the current four are worked out: the
         * loads (L1 or L2, a few hundred cycles) then overlap the arithmetic instead of heading each step. */
        double tvc[4][VS_ITEM_SAMPLES / 32], tvn[4][VS_ITEM_SAMPLES / 32];
        auto fetch = [&](int b, double (&tv)[4][VS_ITEM_SAMPLES / 32]) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int it = items[b + r];
                const double *tp = segs[it & 127].tab + ((it >> 7) * VS_ITEM_SAMPLES + lane);
#pragma unroll
                for (int u = 0; u < VS_ITEM_SAMPLES / 32; u++) tv[r][u] = __ldg(tp + 32 * u);   /* past the segment: some other table entry, unused */
            }
        };
        auto work = [&](int b, const double (&tv)[4][VS_ITEM_SAMPLES / 32]) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int it = items[b + r];
                const VsSeg sg = segs[it & 127];                          /* row*VS_MAXSEG + segment */
                const int nn = (int)(sg.nr & 0xffffu), nrise = (int)(sg.nr >> 16);
                const int k0 = (it >> 7) * VS_ITEM_SAMPLES + lane;
#pragma unroll
                for (int u = 0; u < VS_ITEM_SAMPLES / 32; u++) {
                    const int k = k0 + 32 * u;
                    /* the factor A is multiplied with: h[i] while rising; on the falling branch (K*c - K) + 1 (:328),
                     * which the table already holds unless some stream of the batch varies its closure speed */
                    double fac = tv[r][u];
                    if (general) {
                        const double fall = __dadd_rn(__dsub_rn(__dmul_rn(sg.Kd, fac), sg.Kd), 1.0);
                        fac = k < nrise ? fac : fall;
                    }
                    /* ceil as a 32-bit integer (host-side bounds keep it far from 2^31).  The reference tests the
                     * value after its (short) cast: above 32767 it wraps negative, i.e. below DC; on the falling
                     * branch the argument only decreases, so everything after the first value below DC is DC
                     * too -- also where the short would have wrapped back above DC */
                    const int v = __double2int_ru(__dmul_rn(sg.Ad, fac));
                    if (k < nn && v >= sg.DCi && v <= 32767) sg.out[k] = (int16_t)v;
                }
            }
        };
        if (MODE != VS_MODE_FLOW) {
            if (total > 0) fetch(0, tvc);
            for (int b = 0; b < total; b += 8) {
                if (b + 4 < total) fetch(b + 4, tvn);
                work(b, tvc);
                if (b + 4 < total) {
                    if (b + 8 < total) fetch(b + 8, tvc);
                    work(b + 4, tvn);
                }
            }
        } else {                                    /* flow-only mode: the registers of the second set would cost the
                                                       second CTA per SM, which is worth more there */
            for (int b = 0; b < total; b += 4) {
                fetch(b, tvc);
                work(b, tvc);
            }
        }

        if (NOISE) {                                                /* :385-406 */
            __syncwarp();
            for (int j = row0; j < 32; j += step) {
                if (!lanes[j].noise) continue;
                const int cnt = nsegs[j];
                if (cnt == 0) continue;
                uint32_t st = rngrow[j * 32 + lane];
                for (int sidx = 0; sidx < cnt; sidx++) {
                    const VsSeg sg = segs[j * VS_MAXSEG + sidx];
                    const uint32_t *x = segx + (j * VS_MAXSEG + sidx) * VS_SEGX;
                    const int T3 = (int)(x[0] & 0xffffu), T4 = (int)(x[0] >> 16);
                    const int ndw = (int)x[1], skip = (int)x[2];
                    /* the period's noise samples inside this window, in draw order: [.., T4) then [T3, ..) */
                    const int n1hi = min(sg.a1, T4), n2lo = max(sg.a0, T3);
                    const int c1 = max(0, n1hi - sg.a0), c2 = max(0, sg.a1 - n2lo);
                    /* the perturbation draws to step over come first; a few of them ride along in the scratch */
                    int lead = 0;
                    if (skip > 0 && skip + c1 + c2 <= VS_DRAW_SCRATCH) lead = skip;
                    else if (skip > 0) st = vs_rng_gen(st, skip, lane, nullptr);
                    if (lead + c1 + c2 > 0) {
                        st = vs_rng_gen(st, lead + c1 + c2, lane, scratch);
                        __syncwarp();
                        const int32_t *d1 = scratch + lead - sg.a0, *d2 = scratch + lead + c1 - n2lo;
                        int16_t *po = sg.out - sg.a0;                      /* the period's sample 0 */
                        /* two samples per lane and step: the chains (load, divide, scale, ceil, add, clip) are
                         * long and independent */
                        for (int i = sg.a0 + lane; i < n1hi; i += 64) {
                            const int i2 = i + 32;
                            const bool two = i2 < n1hi;
                            const int wa = vs_noise_w(d1[i], ndw), wb = vs_noise_w(d1[two ? i2 : i], ndw);
                            po[i] = (int16_t)vs_add_clip(po[i], wa);
                            if (two) po[i2] = (int16_t)vs_add_clip(po[i2], wb);
                        }
                        for (int i = n2lo + lane; i < sg.a1; i += 64) {
                            const int i2 = i + 32;
                            const bool two = i2 < sg.a1;
                            const int wa = vs_noise_w(d2[i], ndw), wb = vs_noise_w(d2[two ? i2 : i], ndw);
                            po[i] = (int16_t)vs_add_clip(po[i], wa);
                            if (two) po[i2] = (int16_t)vs_add_clip(po[i2], wb);
                        }
                        __syncwarp();
                    }
                }
                rngrow[j * 32 + lane] = st;
            }
        }
        __syncwarp();
    }
}

/* ---- W: one row, whole 16-byte pieces, consecutive lanes on consecutive pieces ----------------------- */
template <int WIN, bool ALIGNED>
__device__ __forceinline__ void vs_write_row(const int16_t *trow, const VsLane &L, int w, int lane)
{
    const int lo = L.lo, hi = L.hi;
    const int wb = L.blk0 + w * WIN;
    if (wb >= hi || wb + WIN <= lo) return;
#pragma unroll
    for (int pc = lane; pc < WIN / 8; pc += 32) {
        const int m0 = wb + 8 * pc;
        if (m0 + 8 > lo && m0 < hi) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(trow) + pc * 4;
            if (m0 >= lo && m0 + 8 <= hi) {
                uint4 v;
                if (ALIGNED) v = *reinterpret_cast<const uint4 *>(src);
                else { v.x = src[0]; v.y = src[1]; v.z = src[2]; v.w = src[3]; }
                *reinterpret_cast<uint4 *>(L.orow + m0) = v;
            } else {
                const int16_t *s16 = reinterpret_cast<const int16_t *>(src);
                for (int k = 0; k < 8; k++)
                    if (m0 + k >= lo && m0 + k < hi) L.orow[m0 + k] = s16[k];
            }
        }
    }
}

/* ---- F: one lane, one row, WIN samples of the order-22 recurrence, in place ---------------------- */
template <int WIN, bool EXACT, bool RAW, bool CHECKED>
__device__ __forceinline__ void vs_filter_window(uint32_t *row32, double (&y)[VS_RING], const double (&cf)[VS_RING],
                                                 double gaind, double pred, double *rrow, int wb, int lo, int hi)
{
#pragma unroll 1
    for (int b = 0; b < WIN / VS_RING; b++) {
        uint32_t *blk32 = row32 + b * (VS_RING / 2);
#pragma unroll
        for (int kk = 0; kk < VS_RING / 2; kk++) {
            const uint32_t pr = blk32[kk];                            /* two int16 flow samples */
            uint32_t outw = 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int k = 2 * kk + h;
                const int xi = h ? (int)(int16_t)(pr >> 16) : (int)(int16_t)(pr & 0xffffu);
                double acc = __dmul_rn((double)xi, gaind);            /* vowel_new.c:266-269 */
                double v;
                if (EXACT) {
#pragma unroll
                    for (int j = 1; j <= VS_ORDER; j++)               /* :279-281, same order */
                        acc = __dsub_rn(acc, __dmul_rn(-cf[j], y[(k + VS_RING - j) % VS_RING]));
                    v = __dsub_rn(acc, __dmul_rn(pred, y[(k + VS_RING - 1) % VS_RING]));   /* :284 */
                } else {
#pragma unroll
                    for (int j = VS_ORDER; j >= 1; j--)               /* oldest tap first; cf = -A[j] */
                        acc = __fma_rn(y[(k + VS_RING - j) % VS_RING], cf[j], acc);
                    v = __fma_rn(-pred, y[(k + VS_RING - 1) % VS_RING], acc);
                }
                y[k] = acc;                                           /* :287-289 (ring) */
                const int qv = EXACT ? vs_round2int(v) : (CHECKED ? vs_quant_fast(v) : vs_quant_fast_nocheck(v));
                if (RAW) {
                    const int m = wb + b * VS_RING + k;
                    if (m >= lo && m < hi) rrow[m] = v;
                }
                outw |= (uint32_t)(uint16_t)qv << (16 * h);
            }
            blk32[kk] = outw;                                         /* in place */
        }
    }
}

/* FLAGS: bit0 EXACT filter, bit1 RAW output, bit2 CHECKED quantiser */
template <int MODE, bool NOISE, int FLAGS>
__global__ void __launch_bounds__(VS_RENDER_THREADS(MODE), (MODE == VS_MODE_FLOW && !NOISE) ? 2 : 1)   /* flow-only: two CTAs per SM */
vs_render_kernel(const VsRenderArgs a)
{
    constexpr bool PAIRED = MODE != VS_MODE_FLOW;
    constexpr bool DRAWS = NOISE && MODE != VS_MODE_FILTER;
    constexpr bool EXACT = (FLAGS & 1) != 0, RAW = (FLAGS & 2) != 0, CHECKED = (FLAGS & 4) != 0;
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr int NTILE = PAIRED ? 2 : 1;
    constexpr int WIN = vs_win(MODE, NOISE), TS = VS_TS_OF(MODE, WIN), TILE_I16 = VS_TILE_I16_OF(MODE, WIN);
    int16_t *s_tiles = reinterpret_cast<int16_t *>(s_raw);
    VsLane *s_lanes = reinterpret_cast<VsLane *>(s_raw + VS_SMEM_TILES(NTILE, WIN));
    VsSeg *s_segs = reinterpret_cast<VsSeg *>(s_raw + VS_SMEM_TILES(NTILE, WIN) + VS_SMEM_LANES);
    int *s_nseg = reinterpret_cast<int *>(s_raw + VS_SMEM_TILES(NTILE, WIN) + VS_SMEM_LANES + VS_SMEM_SEGS);
    unsigned char *s_items = s_raw + VS_SMEM_TILES(NTILE, WIN) + VS_SMEM_LANES + VS_SMEM_SEGS + VS_SMEM_NSEG;
    uint32_t *s_rng = reinterpret_cast<uint32_t *>(s_raw + VS_SMEM_BASE(NTILE, WIN));
    int32_t *s_scratch = reinterpret_cast<int32_t *>(s_raw + VS_SMEM_BASE(NTILE, WIN) + VS_NP * 32 * 32 * 4);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = warp % VS_NP;                       /* warps 0..NP-1 consume, NP.. produce */
    constexpr int NPROD = VS_PW;                         /* producer warps per group */
    const int prod = PAIRED ? warp / VS_NP - 1 : warp / VS_NP;   /* producer index inside the group, -1 = consumer */
    const bool consumer = PAIRED && warp < VS_NP;
    const int step = NPROD;                              /* a producer warp works rows prod, prod+step, ... */
    int16_t *tile0 = s_tiles + pair * NTILE * TILE_I16;
    VsLane *lanes = s_lanes + pair * 32;
    VsSeg *segs = s_segs + pair * 32 * VS_MAXSEG;
    int *nsegs = s_nseg + pair * (32 + 32 * VS_MAXSEG * VS_SEGX);
    uint32_t *rngrow = DRAWS ? s_rng + pair * 32 * 32 : nullptr;
    int32_t *scratch = DRAWS ? s_scratch + (pair * VS_PW + (prod > 0 ? prod : 0)) * VS_DRAW_SCRATCH : nullptr;
    uint16_t *items = reinterpret_cast<uint16_t *>(s_items + (pair * NPROD + (prod > 0 ? prod : 0)) * VS_ITEMS_BYTES);
    const int group_threads = (1 + VS_PW) * 32;

    /* Row of this lane.  Consumer lane l filters row l; producer lane l keeps the books of row
     * prod + l*step. */
    const int myrow = consumer ? lane : prod + lane * step;
    const uint32_t t = blockIdx.x * VS_NT + pair * 32 + (uint32_t)myrow;
    uint32_t chunk_id = (myrow < 32 && t < a.n_rows) ? a.order[t] : VS_NO_CHUNK;
    const bool active = chunk_id != VS_NO_CHUNK;

    uint32_t q = 0;
    double gaind = 0.0, pred = 0.0;
    double *rrow = nullptr;
    VsLane me;
    me.tab = nullptr; me.ct = nullptr; me.orow = nullptr; me.fin = nullptr;
    me.nstart = 0; me.lo = 0; me.hi = 0; me.blk0 = 0; me.T2 = 0; me.DCi = 0; me.DCs = 0; me.noise = 0; me.tab_cap = 0;
    me.chunk = chunk_id; me.q0 = 0; me.pad = 0;
    if (active) {
        const VsChunk ck = a.chunks[chunk_id];
        const VsStream st = a.streams[ck.stream];
        me.orow = a.pcm_out + st.out_off;
        me.lo = (int)ck.emit_lo; me.hi = (int)ck.emit_hi;
        if (MODE == VS_MODE_FILTER) {
            me.fin = a.flow_in + st.in_off;
            me.nstart = (int)ck.gen_target;
        } else {
            me.tab = a.table + st.tab_off;
            me.ct = a.costab + st.pulse_off;
            me.T2 = st.T2;
            me.DCi = (int)ceilf(st.DC);
            me.DCs = st.DCs;
            me.noise = (st.flags & VS_F_NOISE) ? 1 : 0;
            me.tab_cap = st.tab_cap;
            q = ck.first_period;
            me.q0 = q;
            if (q >= st.tab_cap) {                    /* the plan kernel did not reach this chunk: refuse to walk garbage */
                atomicExch(a.status, VS_ECUDA);
                q = 0; me.hi = 0; me.lo = 0;
            }
            me.nstart = (int)__ldg(&me.tab[q].start);
        }
        const int phase = (int)((reinterpret_cast<uintptr_t>(me.orow) >> 1) & 7);
        me.blk0 = me.nstart - ((phase + me.nstart) & 7);
        gaind = (double)st.gain; pred = (double)st.pre;
        rrow = (RAW && a.raw_out) ? a.raw_out + st.out_off : nullptr;
    }
    /* windows to run: max over the 32 rows of the group (each warp of the group sees a subset) */
    int nwin = active ? (me.hi - me.blk0 + WIN - 1) / WIN : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nwin = max(nwin, __shfl_xor_sync(VS_FULL, nwin, o));
    if (!consumer && myrow < 32) lanes[myrow] = me;          /* producers publish the row descriptors */

    /* noise: the rows' RNG states (oldest word in lane 0), restored from the plan kernel's snapshots */
    if (DRAWS && !consumer) {
        __syncwarp();
        for (int j = prod > 0 ? prod : 0; j < 32; j += step)
            if (lanes[j].noise && lanes[j].chunk != VS_NO_CHUNK)
                rngrow[j * 32 + lane] =                  /* snapshot has f = 3: word 3 is the oldest, words 0..2 the newest */
                    lane < VS_RNG_DEG ? __ldg(a.rng_snap + (size_t)lanes[j].chunk * 32 + (lane + 3) % VS_RNG_DEG) : 0u;
        __syncwarp();
    }

    if (!PAIRED) {
        /* ======== flow mode: every warp does G then W for its own rows ======== */
        __syncwarp();
        for (int w = 0; w < nwin; w++) {
            vs_gen_tile<MODE, NOISE>(tile0, lanes, segs, nsegs, items, w, lane, prod, step, me, q, rngrow, scratch, a.general_pulse != 0);
            __syncwarp();
            for (int j = prod; j < 32; j += step) vs_write_row<WIN, MODE == VS_MODE_FLOW>(tile0 + j * TS, lanes[j], w, lane);
            __syncwarp();
        }
        return;
    }

    /* ======== filtering modes: consumer F(w) overlaps producers' W(w-1) and G(w+1) ======== */
    {
        /* the group's window count: consumers saw all 32 rows, producers only theirs -> share the max */
        __shared__ int s_groupwin[VS_NP];
        if (consumer && lane == 0) s_groupwin[pair] = nwin;
        vs_named_barrier(1 + pair, group_threads);           /* row descriptors + window count visible */
        nwin = s_groupwin[pair];
    }
    if (consumer) {
        double y[VS_RING];
#pragma unroll
        for (int j = 0; j < VS_RING; j++) y[j] = 0.0;

        vs_named_barrier(1 + pair, group_threads);           /* window 0 generated */
        for (int w = 0; w < nwin; w++) {
            const int wb = me.blk0 + w * WIN;
            if (active && wb < me.hi)
                vs_filter_window<WIN, EXACT, RAW, CHECKED>(
                    reinterpret_cast<uint32_t *>(tile0 + (w & 1) * TILE_I16 + lane * TS), y, a.ncf, gaind, pred, rrow,
                    wb, me.lo, me.hi);
            vs_named_barrier(1 + pair, group_threads);
        }
    } else {
        vs_gen_tile<MODE, NOISE>(tile0, lanes, segs, nsegs, items, 0, lane, prod, step, me, q, rngrow, scratch, a.general_pulse != 0);
        vs_named_barrier(1 + pair, group_threads);           /* window 0 generated */
        for (int w = 0; w < nwin; w++) {
            int16_t *other = tile0 + ((w + 1) & 1) * TILE_I16;
            if (w > 0)
                for (int j = prod; j < 32; j += step) vs_write_row<WIN, MODE == VS_MODE_FLOW>(other + j * TS, lanes[j], w - 1, lane);
            if (w + 1 < nwin) {
                vs_gen_tile<MODE, NOISE>(other, lanes, segs, nsegs, items, w + 1, lane, prod, step, me, q, rngrow, scratch, a.general_pulse != 0);
            }
            vs_named_barrier(1 + pair, group_threads);
        }
        if (nwin > 0) {
            const int16_t *last = tile0 + ((nwin - 1) & 1) * TILE_I16;
            for (int j = prod; j < 32; j += step) vs_write_row<WIN, MODE == VS_MODE_FLOW>(last + j * TS, lanes[j], nwin - 1, lane);
        }
    }
}

/* ================================================================================================
 * vowel -n: output noise in place (vowel_new.c:302-324), one WARP per stream.
 * Per frame of `frame` samples: the float power sum runs in sample order (not associative: every lane
 * adds the same 32 squares, staged through shared memory); the frame's random() values come 31 at a time
 * (vs_rng_gen) into a per-warp scratch and the noise is added lane-parallel.
 * ============================================================================================== */
#define VS_VN_NT      128
#define VS_VN_SCRATCH 1024            /* random() values per piece of a frame */

__global__ void __launch_bounds__(VS_VN_NT) vs_vnoise_kernel(int16_t *pcm, const VsNoiseRow *rows, uint32_t n_rows)
{
    __shared__ int32_t s_draws[VS_VN_NT / 32][VS_VN_SCRATCH];
    __shared__ __align__(16) float s_sq[VS_VN_NT / 32][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t s = blockIdx.x * (VS_VN_NT / 32) + wib;
    if (s >= n_rows) return;
    const VsNoiseRow r = rows[s];
    if (!(r.snr > 0.0f) || r.frame == 0) return;
    int32_t *draws = s_draws[wib];
    float *sqb = s_sq[wib];
    uint32_t st = vs_rng_seed_warp(r.seed, lane);                        /* :234 */
    int16_t *y = pcm + r.off;
    for (uint32_t base = 0; base < r.n; base += r.frame) {
        const uint32_t ni = min(r.frame, r.n - base);
        float aux = 0.0f;
        for (uint32_t b0 = 0; b0 < ni; b0 += 32) {                       /* :304-306, float, in order */
            const uint32_t i = b0 + lane;
            const float v = i < ni ? (float)y[base + i] : 0.0f;
            sqb[lane] = __fmul_rn(v, v);                                 /* +0.0f beyond the frame: adds nothing */
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 q = reinterpret_cast<const float4 *>(sqb)[k];
                aux = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(aux, q.x), q.y), q.z), q.w);
            }
            __syncwarp();
        }
        const float sig_power = __fdiv_rn(aux, (float)(int16_t)ni);      /* `ni` is a signed short (:65) */
        const float width = __double2float_rn(sqrt((double)__fdiv_rn(__fmul_rn(12.0f, sig_power), r.snr)));   /* :309 */
        for (uint32_t p0 = 0; p0 < ni; p0 += VS_VN_SCRATCH) {            /* :314-319 */
            const uint32_t m = min((uint32_t)VS_VN_SCRATCH, ni - p0);
            st = vs_rng_gen(st, (int)m, lane, draws);
            __syncwarp();
            for (uint32_t k = lane; k < m; k += 32) {
                const float nv = __double2float_rn(vs_div_const((double)draws[k], VS_RAND_MAX_D, VS_INV_RM));
                const float a = __double2float_rn(__dmul_rn((double)width, __dsub_rn((double)nv, 0.5)));
                int16_t *dst = y + base + p0 + k;
                *dst = (int16_t)vs_round2int(__dadd_rn((double)*dst, (double)a));
            }
            __syncwarp();
        }
    }
}

cudaError_t vs_launch_vnoise(int16_t *pcm, const VsNoiseRow *rows, uint32_t n_rows, cudaStream_t s)
{
    vs_vnoise_kernel<<<(n_rows + VS_VN_NT / 32 - 1) / (VS_VN_NT / 32), VS_VN_NT, 0, s>>>(pcm, rows, n_rows);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------------
 * FP64 pipe peak: 8 independent DFMA chains per thread, every SM full.  Used by bench.py to put a
 * measured denominator next to the HBM-write roofline.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) vs_fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t vs_launch_fp64_peak(double *scratch, int blocks, int iters, cudaStream_t s)
{
    vs_fp64_peak_kernel<<<blocks, 256, 0, s>>>(scratch, iters, 0.999999, 1e-9);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------------
 * launch wrappers (called from vs_api.cu)
 * ---------------------------------------------------------------------------------------------- */
cudaError_t vs_launch_plan(const VsPlanArgs &a, bool want_log, bool warp_per_stream, cudaStream_t s)
{
    if (warp_per_stream && !want_log) {
        vs_plan_warp_kernel<<<(a.n_streams + VS_PLANW_NT / 32 - 1) / (VS_PLANW_NT / 32), VS_PLANW_NT, 0, s>>>(a);
        return cudaGetLastError();
    }
    const unsigned grid = (a.n_streams + VS_PLAN_NT - 1) / VS_PLAN_NT;
    static_assert(VS_PLAN_SMEM >= VS_RNG_DEG * VS_PLAN_NT * 4, "plan kernel shared memory");
    if (want_log) {
        cudaFuncSetAttribute(vs_plan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VS_PLAN_SMEM);
        vs_plan_kernel<true><<<grid, VS_PLAN_NT, VS_PLAN_SMEM, s>>>(a);
    } else {
        cudaFuncSetAttribute(vs_plan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VS_PLAN_SMEM);
        vs_plan_kernel<false><<<grid, VS_PLAN_NT, VS_PLAN_SMEM, s>>>(a);
    }
    return cudaGetLastError();
}

template <int MODE, bool NOISE, int FLAGS>
static void vs_go(const VsRenderArgs &a, cudaStream_t s)
{
    const unsigned grid = a.n_rows / VS_NT;
    const int dyn = VS_SMEM_BASE(MODE == VS_MODE_FLOW ? 1 : 2, vs_win(MODE, NOISE)) + ((NOISE && MODE != VS_MODE_FILTER) ? VS_SMEM_NOISE : 0);
    cudaFuncSetAttribute(vs_render_kernel<MODE, NOISE, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    vs_render_kernel<MODE, NOISE, FLAGS><<<grid, VS_RENDER_THREADS(MODE), dyn, s>>>(a);
}

template <int MODE, bool NOISE>
static void vs_go_flags(const VsRenderArgs &a, bool exact, cudaStream_t s)
{
    const bool raw = a.raw_out != nullptr;
    if (exact) { if (raw) vs_go<MODE, NOISE, 3>(a, s); else vs_go<MODE, NOISE, 1>(a, s); }
    else if (raw) vs_go<MODE, NOISE, 6>(a, s);                       /* raw output is a debug path: keep it checked */
    else if (a.checked_quant) vs_go<MODE, NOISE, 4>(a, s);
    else vs_go<MODE, NOISE, 0>(a, s);
}

cudaError_t vs_launch_render(const VsRenderArgs &a, int mode, bool noise, bool exact, cudaStream_t s)
{
    if (mode == VS_MODE_FLOW) { if (noise) vs_go<VS_MODE_FLOW, true, 0>(a, s); else vs_go<VS_MODE_FLOW, false, 0>(a, s); }
    else if (mode == VS_MODE_FILTER) vs_go_flags<VS_MODE_FILTER, false>(a, exact, s);
    else if (noise) vs_go_flags<VS_MODE_SYNTH, true>(a, exact, s);
    else vs_go_flags<VS_MODE_SYNTH, false>(a, exact, s);
    return cudaGetLastError();
}
