/* vs_render.cu -- the sm_100a render kernel of libvoicesynth_cuda: ONE THREAD PER STREAM-CHUNK.
 *
 * A CTA is 4 warps (one per SM sub-partition); lane l of a warp owns one row = one time-chunk of one voice
 * and walks down it sample by sample, all 32 lanes in lock step (SIMT across voices):
 *
 *   G  generate   the glottal flow sample (flowgen_shimmer.c:316-336, noise :385-406) from the period table the
 *                 plan kernel wrote: x = ceil(A * table[i]) inside the open phase, DC outside.  Branch free:
 *                 8 samples at a time, the lane's current and next pitch period both in registers, a sample
 *                 picks its period by a bit of a per-group mask (GEN_FAST).  Period entries reach the lane
 *                 through a small per-lane ring in shared memory that cp.async refills one window ahead;
 *                 pulse tables are staged in shared memory once per warp.
 *   F  filter     the order-22 all-pole recurrence (vowel_new.c:266-289) in FP64 registers: a 24-entry ring,
 *                 fully unrolled, coefficients in UNIFORM registers (the vowel preset is a function of
 *                 blockIdx and kernel parameters).  23 FP64-pipe instructions per sample when gain is integral and
 *                 pre-emphasis 0 or 1 (both commute to the integer input: x' = gain*(x[n] - pre*x[n-1])),
 *                 25 otherwise; the FP64 pipe is the bound.  Quantiser: one F2I + clamp.
 *   W  write      every lane packs 8 results into one 16-byte shared-memory store down its own tile row
 *                 (row stride 16 x odd bytes: conflict free); once per window every lane hands its row
 *                 to the TMA engine as one bulk shared->global copy (cp.async.bulk, SASS UBLKCP) -- whole
 *                 16-byte pieces, 16-byte aligned in HBM -- and carries on with the other tile.
 *
 * The flow never exists in memory: in the fused mode only int16 PCM leaves the SM.
 * GEN_SIMPLE is the general generator (per-sample, branching): glottal noise, -z, very short pitch periods.
 * Exact mode (VS_OPT_EXACT_FILTER) keeps the reference's unfused multiply/subtract order and its floor-based
 * round2int() bit for bit.
 */
#include "vs_device.cuh"
#include "vs_presets.h"

static_assert(VS_NUM_PRESETS == VS_NUM_PRESETS_I, "preset count");

__constant__ double c_ncoef[VS_NUM_PRESETS][VS_RING];      /* c_ncoef[p][j] = -A_p[j], j = 1..22 (vowel_new.c:450-544) */

enum { VS_MODE_FLOW = 0, VS_MODE_SYNTH = 1, VS_MODE_FILTER = 2 };
enum { VS_GEN_FAST = 0, VS_GEN_SIMPLE = 1 };
enum { VS_FILT_INT = 0, VS_FILT_FMA = 1, VS_FILT_EXACT = 2 };

__host__ __device__ constexpr int vs_win(int mode) { return mode == VS_MODE_SYNTH ? VS_WIN_SYNTH : (mode == VS_MODE_FLOW ? VS_WIN_FLOW : VS_WIN_FILTER); }
#define VS_BIG_T 0x3fffffff
#define VS_GROUP 8

/* ---- PTX: async copies -------------------------------------------------------------------------------- */
__device__ __forceinline__ void vs_cp_async8(uint32_t dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void vs_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
/* this thread's shared-memory writes become visible to the async proxy (the TMA engine) */
__device__ __forceinline__ void vs_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void vs_bulk_s2g(void *gdst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vs_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void vs_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

/* the glottal-noise sample (flowgen_shimmer.c:387,394) in two FP64 operations: t = r/M - 1/2, w = ceil(t*NDW).
 * Equal to the reference's divide / multiply / subtract / ceil for every r and NDW < 2^19 (proof and exhaustive
 * check around every breakpoint: tests/tools/noisecheck.c). */
__device__ __forceinline__ int vs_noise_w2(int32_t r, double ndwd)
{
    double t = __fma_rn((double)r, VS_INV_RM, -0.5);
    if (r == 2147483647) t = 0.5;
    return vs_ceil_s16(__dmul_rn(t, ndwd));
}

template <int MODE, int GEN, bool NOISE, int FILT, bool RAW>
__global__ void __launch_bounds__(VS_NT, MODE == VS_MODE_FLOW ? 3 : 1) vs_render_kernel(const VsRenderArgs a)
{
    constexpr int WIN = vs_win(MODE), TSB = WIN * 2, TILE = 32 * TSB;
    constexpr bool HASGEN = MODE != VS_MODE_FILTER, HASFILT = MODE != VS_MODE_FLOW;
    constexpr bool FAST = HASGEN && GEN == VS_GEN_FAST;
    static_assert((WIN % VS_RING) == 0 && ((WIN / VS_RING) & 1) == 1, "window = odd number of ring blocks");
    static_assert(!(FAST && NOISE), "the fast generator has no noise path");
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    /* vowel preset of this CTA: uniform by construction */
    int preset = 0;
    if (HASFILT) {
#pragma unroll
        for (int p = 0; p < VS_NUM_PRESETS - 1; p++) preset += blockIdx.x >= a.cta_end[p] ? 1 : 0;
    }

    /* shared memory: per warp [tile 0 | tile 1 | period ring | pulse-table cache], then the CTA's RNG states */
    const uint32_t tile_off = (uint32_t)warp * a.warp_bytes;
    const uint32_t ring_off = tile_off + 2u * TILE;
    const uint32_t cache_off = ring_off + a.ring_R * 256u;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t *s_rng = reinterpret_cast<uint32_t *>(smem + 4u * a.warp_bytes);      /* [31][VS_NT], NOISE only */

    /* ---- the lane's row ------------------------------------------------------------------------------- */
    const uint32_t t = blockIdx.x * VS_NT + threadIdx.x;
    const uint32_t chunk_id = t < a.n_rows ? __ldg(a.order + t) : VS_NO_CHUNK;
    const bool active = chunk_id != VS_NO_CHUNK;
    int lo = 0, hi = 0, nstart = 0;
    int16_t *orow = nullptr;
    const int16_t *fin = nullptr;
    double *rrow = nullptr;
    double gaind = 0.0, pred = 0.0;
    int gain_i = 0, pre_i = 0;
    int T2 = 0, DCi = 0, DCs = 0, q0 = 0, nper = 0;
    uint32_t pulse_off = 0xffffffffu;
    bool nz = false;
    const unsigned char *ptab = nullptr;                    /* the row's period table */
    if (active) {
        const VsChunk ck = a.chunks[chunk_id];
        const VsStream *st = a.streams + ck.stream;
        orow = a.pcm_out + st->out_off;
        lo = (int)ck.emit_lo; hi = (int)ck.emit_hi;
        gaind = (double)st->gain; pred = (double)st->pre;
        gain_i = (int)st->gain; pre_i = (int)st->pre;
        rrow = (RAW && a.raw_out) ? a.raw_out + st->out_off : nullptr;
        if (MODE == VS_MODE_FILTER) {
            fin = a.flow_in + st->in_off;
            nstart = (int)ck.gen_target;
        } else {
            ptab = reinterpret_cast<const unsigned char *>(a.table) + st->tab_off * (a.compact ? sizeof(VsPeriodC) : sizeof(VsPeriod));
            T2 = st->T2;
            DCi = (int)ceilf(st->DC);                       /* (float)x < DC  <=>  x < ceil(DC) for integer x */
            DCs = st->DCs;
            nz = (st->flags & VS_F_NOISE) != 0;
            pulse_off = st->pulse_off;
            q0 = (int)ck.first_period;
            nper = (int)__ldg(a.n_periods + ck.stream);
            nstart = (int)ck.first_start;
            if (q0 >= nper || q0 >= (int)st->tab_cap) {     /* the plan kernel did not reach this chunk: refuse to walk garbage */
                atomicExch(a.status, VS_ECUDA);
                nper = 0; hi = 0; lo = 0; nstart = 0;
            }
        }
    }
    /* windows are anchored on the 16-byte grid of the row's own address in HBM */
    const int phase = (int)((reinterpret_cast<uintptr_t>(orow) >> 1) & 7);
    const int blk0 = nstart - ((phase + nstart) & 7);
    int nwin = active ? (hi - blk0 + WIN - 1) / WIN : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nwin = max(nwin, __shfl_xor_sync(VS_FULL, nwin, o));

    /* ---- filter state ------------------------------------------------------------------------------------ */
    double y[VS_RING], cf[VS_RING];
#pragma unroll
    for (int j = 0; j < VS_RING; j++) { y[j] = 0.0; cf[j] = HASFILT ? c_ncoef[preset][j] : 0.0; }
    int xg_prev = 0;

    /* ---- generator state -------------------------------------------------------------------------------- */
    const int qlast = nper - 1;
    int q = q0 - 1;                                         /* index of the current period                         */
    /* fast generator: cur = (Adc, Tc, noc), next = (An, Tn, non); ic = in-period index of the group's first sample */
    int ic = 0, Tc = nstart - blk0, noc = 0, Tn = VS_BIG_T, non = 0;
    double Adc = 0.0;
    float An = 0.0f;
    uint32_t tb = cache_off;                                /* the row's pulse table in shared memory (byte offset) */
    uint32_t qf = (uint32_t)q0;                             /* first period entry not yet requested                */
    const uint32_t Rm = a.ring_R - 1u;
    /* simple generator */
    int si = 0, sT = nstart - blk0, snopen = 0, sT3 = VS_BIG_T, sT4 = 0, snpert = 0, dcs_cur = 0;   /* no noise in the slots before the first period */
    double sAd = 0.0, sKd = 0.0, sndwd = 0.0;
    const double *gtab = a.costab + (pulse_off == 0xffffffffu ? 0u : pulse_off);
    VsRng rng;
    rng.r = s_rng + threadIdx.x;
    rng.f = 3;

    auto nopen_of = [&](float A, int T) -> int { return min(T, A > 32767.0f ? T2 : 2 * T2); };

    if (FAST) {
        /* pulse tables into the warp's cache: one copy per distinct table of the warp's rows (the host sized the
         * cache from the same row order, so everything fits) */
        const uint32_t key = active ? pulse_off : 0xffffffffu;
        const uint32_t grp = __match_any_sync(VS_FULL, key);
        const int leader = __ffs((int)grp) - 1;
        const uint32_t len = active ? 2u * (uint32_t)T2 : 0u;
        const uint32_t mine = (active && lane == leader) ? len : 0u;
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(VS_FULL, incl, o);
            if (lane >= o) incl += up;
        }
        const uint32_t base = incl - mine;
        const uint32_t lbase = __shfl_sync(VS_FULL, base, leader);
        uint32_t leaders = __ballot_sync(VS_FULL, mine > 0u);
        double *cache = reinterpret_cast<double *>(smem + cache_off);
        while (leaders) {
            const int L = __ffs((int)leaders) - 1;
            leaders &= leaders - 1u;
            const uint32_t src = __shfl_sync(VS_FULL, pulse_off, L), n = __shfl_sync(VS_FULL, len, L), dst = __shfl_sync(VS_FULL, base, L);
            if (dst + n <= a.cache_doubles)
                for (uint32_t k = lane; k < n; k += 32) cache[dst + k] = __ldg(a.costab + src + k);
            else if (lane == 0) atomicExch(a.status, VS_ECUDA);
        }
        tb = cache_off + lbase * 8u;
        /* first entries of the period ring, then the first period as `next` of an empty current period that
         * covers the slots before the row's first sample */
        if (active) {
            const uint32_t want = min((uint32_t)q0 + a.ring_ahead, (uint32_t)nper);
            for (; qf < want; qf++)
                vs_cp_async8(smem_base + ring_off + ((qf & Rm) * 32u + lane) * 8u, ptab + (size_t)qf * sizeof(VsPeriodC));
        }
        vs_cp_async_wait_all();
        __syncwarp();
        if (active && q0 <= qlast) {
            const uint2 e = *reinterpret_cast<const uint2 *>(smem + ring_off + (((uint32_t)q0 & Rm) * 32u + lane) * 8u);
            An = __uint_as_float(e.x); Tn = (int)e.y;
            non = nopen_of(An, Tn);
        }
    }
    if (NOISE && HASGEN && active && nz && nper > 0) {
        /* the plan kernel's snapshot is taken after the first period's K draw; it is stored with f = 3 */
        for (int k = 0; k < VS_RNG_DEG; k++) rng.r[k * VS_NT] = __ldg(a.rng_snap + (size_t)chunk_id * 32 + k);
    }

    /* ---- G, fast: 8 samples of the lane's row, branch free -------------------------------------------- */
    auto gen_fast = [&](int (&x)[VS_GROUP], const bool first) {
        const int tn = Tc - ic;                             /* samples of the current period left             */
        const int kb = min(tn, VS_GROUP);                   /* sample u >= kb belongs to the next period      */
        const int oc = max(min(noc - ic, kb), 0);           /* open-phase samples of the current period       */
        const int on = min(kb + non, VS_GROUP);             /* the next period is open on [kb, on)            */
        const uint32_t selm = 0xffu << kb;
        const uint32_t openm = ((1u << oc) - 1u) | (selm & ((1u << on) - 1u));
        const uint32_t a_c = tb + (uint32_t)ic * 8u, a_n = tb - (uint32_t)kb * 8u;
        const double Adn = (double)An;
#pragma unroll
        for (int u = 0; u < VS_GROUP; u++) {
            const bool ps = ((selm >> u) & 1u) != 0u, po = ((openm >> u) & 1u) != 0u;
            const uint32_t ad = ps ? a_n : a_c;
            double fac = 0.0;
            if (po) fac = *reinterpret_cast<const double *>(smem + ad + u * 8);
            const double A = ps ? Adn : Adc;
            /* ceil as a 32-bit integer.  The reference tests the value after its (short) cast: above 32767 it wraps
             * negative, i.e. below DC (flowgen_shimmer.c:320,329) */
            const int v = __double2int_ru(__dmul_rn(A, fac));
            const bool ok = po && v >= DCi && v <= 32767;
            x[u] = ok ? v : (first ? (ps ? DCs : 0) : DCs);
        }
        /* the group's last sample may have been the period's last: promote `next`, read the entry after it */
        const bool pr = tn <= VS_GROUP;
        ic += VS_GROUP;
        if (pr) { ic -= Tc; Adc = Adn; noc = non; Tc = Tn; q++; }
        const uint2 e = *reinterpret_cast<const uint2 *>(smem + ring_off + ((((uint32_t)(q + 1)) & Rm) * 32u + lane) * 8u);
        const bool valid = q < qlast;
        An = valid ? __uint_as_float(e.x) : 0.0f;
        Tn = valid ? (int)e.y : VS_BIG_T;
        non = nopen_of(An, Tn);
    };

    /* keep the ring `ring_ahead` periods ahead of the current one (called once per window) */
    auto ring_refill = [&]() {
        vs_cp_async_wait_all();                             /* what the previous window requested has long landed */
        const uint32_t want = min((uint32_t)(q + 1) + a.ring_ahead, (uint32_t)max(nper, 0));
        for (uint32_t k = 0; k < a.ring_fetch; k++) {
            const uint32_t idx = qf + k;
            if (idx < want) vs_cp_async8(smem_base + ring_off + ((idx & Rm) * 32u + lane) * 8u, ptab + (size_t)idx * sizeof(VsPeriodC));
        }
        qf = max(qf, min(want, qf + a.ring_fetch));
    };

    /* ---- G, simple: one sample (any parameters: glottal noise, -z, short periods) ---------------------- */
    auto gen_simple = [&]() -> int {
        while (si >= sT) {                                  /* next pitch period */
            q++;
            si = 0;
            if (q > qlast) { sT = VS_BIG_T; snopen = 0; sAd = 0.0; break; }
            dcs_cur = DCs;
            float A;
            if (a.compact) {
                const uint2 e = __ldg(reinterpret_cast<const uint2 *>(ptab) + q);           /* VsPeriodC */
                A = __uint_as_float(e.x); sT = (int)e.y;
            } else {
                const VsPeriod *e = reinterpret_cast<const VsPeriod *>(ptab) + q;
                const double2 ak = __ldg(reinterpret_cast<const double2 *>(e));
                const int4 b = __ldg(reinterpret_cast<const int4 *>(e) + 1);
                A = (float)ak.x; sKd = ak.y;
                sT = b.y & 0xffff; snpert = (int)((uint32_t)b.y >> 16);
                sT3 = b.z & 0xffff; sT4 = (int)((uint32_t)b.z >> 16);
                sndwd = (double)b.w;
                if (NOISE && nz && q != q0)                 /* the period's jitter / shimmer / K draws (:283,:298,:325) */
                    for (int k = 0; k < snpert; k++) (void)vs_rng_next<VS_NT>(rng);
            }
            sAd = (double)A;
            snopen = nopen_of(A, sT);                       /* A above 32767: x[T2] = (short)ceil(A) < 0, the falling branch is left at once */
        }
        int x = dcs_cur;
        if (si < snopen) {
            double fac = __ldg(gtab + si);
            if (a.general_pulse && si >= T2) fac = __dadd_rn(__dsub_rn(__dmul_rn(sKd, fac), sKd), 1.0);      /* :328 */
            const int v = __double2int_ru(__dmul_rn(sAd, fac));
            if (v >= DCi && v <= 32767) x = v;
        }
        if (NOISE && nz && sT != VS_BIG_T && (si < sT4 || si >= sT3)) {                                      /* :385-399 */
            const int32_t r = vs_rng_next<VS_NT>(rng);
            x = vs_add_clip(x, vs_noise_w2(r, sndwd));
        }
        si++;
        return x;
    };

    int mbase = blk0;                                       /* stream index of the next sample to filter / store */
    auto gen_group = [&](int (&x)[VS_GROUP], const bool first, const int m0) {
        if (MODE == VS_MODE_FILTER) {
#pragma unroll
            for (int u = 0; u < VS_GROUP; u++) {
                const int m = m0 + u;
                x[u] = (m >= nstart && m < hi) ? (int)__ldg(fin + m) : 0;
            }
        } else if (FAST) {
            gen_fast(x, first);
        } else {
#pragma unroll 1
            for (int u = 0; u < VS_GROUP; u++) x[u] = gen_simple();
        }
    };

    /* ---- F: 8 samples of the recurrence at ring positions k0.., packed into one 16-byte piece ---------- */
    auto quant = [&](double v) -> int {
        if (FILT == VS_FILT_EXACT) return vs_round2int(v);                       /* vowel_new.c:413-427, literally */
        return max(-32767, min(32767, __double2int_rn(v)));                     /* F2I saturates; ties (never hit) go to even */
    };

    /* ---- W: the finished window leaves through the TMA engine ------------------------------------------ */
    auto store_window = [&](const int w) {
        const int wb = blk0 + w * WIN;
        const int a0 = max(wb, lo), b0 = min(wb + WIN, hi);
        vs_fence_async();
        if (b0 > a0) {
            const uint32_t trow = tile_off + (uint32_t)(w & 1) * TILE + (uint32_t)lane * TSB;
            const int a8 = wb + ((a0 - wb + 7) & ~7), b8 = wb + ((b0 - wb) & ~7);
            if (b8 > a8) vs_bulk_s2g(orow + a8, smem_base + trow + (uint32_t)(a8 - wb) * 2u, (uint32_t)(b8 - a8) * 2u);
            const int16_t *tr = reinterpret_cast<const int16_t *>(smem + trow);
            const int hend = min(a8, b0);
            for (int m = a0; m < hend; m++) orow[m] = tr[m - wb];               /* a stream's first and last few samples */
            for (int m = max(b8, hend); m < b0; m++) orow[m] = tr[m - wb];
        }
        vs_bulk_commit();
    };

    /* ======== the row, window by window ======== */
    int xn[VS_GROUP];
    gen_group(xn, true, blk0);
    for (int w = 0; w < nwin; w++) {
        if (FAST) ring_refill();
        if (w >= 2) vs_bulk_wait_read<1>();                 /* the copy that last read this tile has finished */
        unsigned char *trow = smem + tile_off + (uint32_t)(w & 1) * TILE + (uint32_t)lane * TSB;
#pragma unroll 1
        for (int b = 0; b < WIN / VS_RING; b++) {
#pragma unroll
            for (int g = 0; g < VS_RING / VS_GROUP; g++) {
                int xc[VS_GROUP];
#pragma unroll
                for (int u = 0; u < VS_GROUP; u++) xc[u] = xn[u];
                gen_group(xn, false, mbase + VS_GROUP);     /* one group ahead: its latency hides behind the filter */
                uint32_t ow[VS_GROUP / 2];
#pragma unroll
                for (int u = 0; u < VS_GROUP; u++) {
                    const int k = g * VS_GROUP + u;
                    int qv;
                    if (!HASFILT) {
                        qv = xc[u];
                    } else {
                        double acc, v;
                        if (FILT == VS_FILT_INT) {
                            /* gain and pre-emphasis on the integer input; the recurrence then yields the
                             * pre-emphasised waveform directly (the filter is LTI) */
                            const int xg = xc[u] * gain_i;
                            acc = (double)(xg - xg_prev * pre_i);
                            xg_prev = xg;
#pragma unroll
                            for (int j = VS_ORDER; j >= 1; j--) acc = __fma_rn(y[(k + VS_RING - j) % VS_RING], cf[j], acc);
                            v = acc;
                        } else if (FILT == VS_FILT_FMA) {
                            acc = __dmul_rn((double)xc[u], gaind);                                   /* vowel_new.c:266-269 */
#pragma unroll
                            for (int j = VS_ORDER; j >= 1; j--) acc = __fma_rn(y[(k + VS_RING - j) % VS_RING], cf[j], acc);
                            v = __fma_rn(-pred, y[(k + VS_RING - 1) % VS_RING], acc);                /* :284 */
                        } else {
                            acc = __dmul_rn((double)xc[u], gaind);
#pragma unroll
                            for (int j = 1; j <= VS_ORDER; j++)                                      /* :279-281, same order, unfused */
                                acc = __dsub_rn(acc, __dmul_rn(-cf[j], y[(k + VS_RING - j) % VS_RING]));
                            v = __dsub_rn(acc, __dmul_rn(pred, y[(k + VS_RING - 1) % VS_RING]));
                        }
                        y[k] = acc;                                                                  /* :287-289 (ring) */
                        qv = quant(v);
                        if (RAW) {
                            const int m = mbase + u;
                            if (rrow && m >= lo && m < hi) rrow[m] = v;
                        }
                    }
                    if (u & 1) ow[u >> 1] |= (uint32_t)qv << 16;
                    else ow[u >> 1] = (uint32_t)qv & 0xffffu;
                }
                *reinterpret_cast<uint4 *>(trow + (b * VS_RING + g * VS_GROUP) * 2) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                mbase += VS_GROUP;
            }
        }
        store_window(w);
    }
    vs_bulk_wait_read<0>();                                 /* shared memory must outlive the copies that read it */
}

/* ------------------------------------------------------------------------------------------------
 * host side of the kernel: coefficient upload, shared-memory size, launch
 * ---------------------------------------------------------------------------------------------- */
cudaError_t vs_render_init_device()
{
    static double h[VS_NUM_PRESETS][VS_RING];
    for (int p = 0; p < VS_NUM_PRESETS; p++)
        for (int j = 0; j < VS_RING; j++) h[p][j] = j <= VS_ORDER ? -vs_preset_den[p][j] : 0.0;
    return cudaMemcpyToSymbol(c_ncoef, h, sizeof h);
}

int vs_render_window(int mode) { return vs_win(mode); }

template <int MODE, int GEN, bool NOISE, int FILT, bool RAW>
static cudaError_t vs_go(const VsRenderArgs &a, cudaStream_t s)
{
    const unsigned grid = a.n_rows / VS_NT;
    const int dyn = 4 * (int)a.warp_bytes + (NOISE ? VS_RNG_DEG * VS_NT * 4 : 0);
    /* the attribute is per device and per kernel: set it whenever a launch needs more than the last one did */
    static int granted[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && granted[dev] < dyn) {
        const cudaError_t e = cudaFuncSetAttribute(vs_render_kernel<MODE, GEN, NOISE, FILT, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        if (e != cudaSuccess) return e;
        granted[dev] = dyn;
    } else if (dev >= 16) {
        cudaFuncSetAttribute(vs_render_kernel<MODE, GEN, NOISE, FILT, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    }
    vs_render_kernel<MODE, GEN, NOISE, FILT, RAW><<<grid, VS_NT, dyn, s>>>(a);
    return cudaGetLastError();
}

template <int MODE, int GEN, bool NOISE>
static cudaError_t vs_go_filt(const VsRenderArgs &a, int filt, cudaStream_t s)
{
    const bool raw = a.raw_out != nullptr;
    if (filt == VS_FILT_EXACT) return raw ? vs_go<MODE, GEN, NOISE, VS_FILT_EXACT, true>(a, s) : vs_go<MODE, GEN, NOISE, VS_FILT_EXACT, false>(a, s);
    if (raw) return vs_go<MODE, GEN, NOISE, VS_FILT_FMA, true>(a, s);
    if (filt == VS_FILT_INT) return vs_go<MODE, GEN, NOISE, VS_FILT_INT, false>(a, s);
    return vs_go<MODE, GEN, NOISE, VS_FILT_FMA, false>(a, s);
}

/* gen: VS_GEN_FAST / VS_GEN_SIMPLE; filt: VS_FILT_INT / _FMA / _EXACT (raw output forces _FMA or _EXACT) */
cudaError_t vs_launch_render(const VsRenderArgs &a, int mode, int gen, bool noise, int filt, cudaStream_t s)
{
    if (mode == VS_MODE_FILTER) return vs_go_filt<VS_MODE_FILTER, VS_GEN_SIMPLE, false>(a, filt, s);
    if (mode == VS_MODE_FLOW) {
        if (gen == VS_GEN_FAST && !noise) return vs_go<VS_MODE_FLOW, VS_GEN_FAST, false, VS_FILT_FMA, false>(a, s);
        return noise ? vs_go<VS_MODE_FLOW, VS_GEN_SIMPLE, true, VS_FILT_FMA, false>(a, s) : vs_go<VS_MODE_FLOW, VS_GEN_SIMPLE, false, VS_FILT_FMA, false>(a, s);
    }
    if (gen == VS_GEN_FAST && !noise) return vs_go_filt<VS_MODE_SYNTH, VS_GEN_FAST, false>(a, filt, s);
    return noise ? vs_go_filt<VS_MODE_SYNTH, VS_GEN_SIMPLE, true>(a, filt, s) : vs_go_filt<VS_MODE_SYNTH, VS_GEN_SIMPLE, false>(a, filt, s);
}
