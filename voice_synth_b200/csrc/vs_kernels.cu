/* vs_kernels.cu -- sm_100a kernels of libvoicesynth_cuda.
 *
 *   vs_plan_kernel    one thread per STREAM, lock-step over pitch periods: the strictly sequential
 *                     part of flowgen_shimmer.c:246-423 -- glibc random() state, jitter and shimmer
 *                     random walks with their rejection loops, the closure-speed draw, and (with
 *                     -n) the pulse power and the number of noise draws.  Emits a period table and,
 *                     per time-chunk, the period the chunk starts in plus the RNG state there.
 *   vs_render_kernel  one thread per (stream, time-chunk), lock-step over samples: pulse samples
 *                     from the period table + host-libm cosine tables, closed-phase noise, and
 *                     (fused / filter modes) the order-22 FP64 all-pole recurrence of
 *                     vowel_new.c:252-296 with its round-half-down quantiser, state kept in a
 *                     24-entry register ring so that one unrolled block yields 3 x 16-byte stores.
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

/* ------------------------------------------------------------------------------------------------
 * glibc random() TYPE_3 (r[i] = r[i-3] + r[i-31], output >> 1).  State lives in shared memory,
 * word-major ([31][VS_NT]) so that lanes never collide on a bank whatever their private index is.
 * ---------------------------------------------------------------------------------------------- */
struct VsRng {
    uint32_t *r;   /* shared memory base + threadIdx.x */
    int f;         /* front index; back index is f-3 (mod 31) */
};

__device__ __forceinline__ int32_t vs_rng_next(VsRng &g)
{
    const int b = g.f >= 3 ? g.f - 3 : g.f + 28;
    const uint32_t v = g.r[g.f * VS_NT] + g.r[b * VS_NT];
    g.r[g.f * VS_NT] = v;
    g.f = (g.f == VS_RNG_DEG - 1) ? 0 : g.f + 1;
    return (int32_t)(v >> 1);
}

__device__ void vs_rng_seed(VsRng &g, uint32_t seed)
{
    int32_t w = (int32_t)(seed ? seed : 1u);
    g.r[0] = (uint32_t)w;
    for (int i = 1; i < VS_RNG_DEG; i++) {          /* 16807*w mod (2^31-1), Schrage, signed */
        const int32_t hi = w / 127773, lo = w % 127773;
        w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        g.r[i * VS_NT] = (uint32_t)w;
    }
    g.f = 3;
    for (int i = 0; i < 310; i++) (void)vs_rng_next(g);
}

/* store / load the state in canonical rotation (oldest word first => f = 3 after loading) */
__device__ void vs_rng_save(const VsRng &g, uint32_t *dst)
{
    int j = g.f >= 3 ? g.f - 3 : g.f + 28;
    for (int k = 0; k < VS_RNG_DEG; k++) {
        dst[k] = g.r[j * VS_NT];
        j = (j == VS_RNG_DEG - 1) ? 0 : j + 1;
    }
}
__device__ void vs_rng_load(VsRng &g, const uint32_t *src)
{
    for (int k = 0; k < VS_RNG_DEG; k++) g.r[k * VS_NT] = src[k];
    g.f = 3;
}

/* ------------------------------------------------------------------------------------------------
 * pulse samples (flowgen_shimmer.c:319, :328) and the noise sample (:387, :394, :591-600)
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int16_t vs_rising(double Ah, double c)
{
    return vs_d2s(ceil(__dmul_rn(Ah, __dsub_rn(1.0, c))));
}
__device__ __forceinline__ int16_t vs_falling(double Ad, double Kd, double c)
{
    return vs_d2s(ceil(__dmul_rn(Ad, __dadd_rn(__dsub_rn(__dmul_rn(Kd, c), Kd), 1.0))));
}
__device__ __forceinline__ int16_t vs_noise_w(int32_t r, int32_t ndw)
{
    const double u = __ddiv_rn((double)r, VS_RAND_MAX_D);
    return vs_d2s(ceil(__dsub_rn(__dmul_rn(u, (double)ndw), __ddiv_rn((double)ndw, 2.0))));
}
__device__ __forceinline__ int16_t vs_clip_ceil(float v)
{
    if (v > 32767.0f) return 32767;
    if (v < -32767.0f) return -32767;
    return vs_d2s(ceil((double)v));
}

/* ================================================================================================
 * PLAN: one thread per stream
 * ============================================================================================== */
template <bool LOG>
__global__ void __launch_bounds__(VS_NT) vs_plan_kernel(const VsPlanArgs a)
{
    __shared__ uint32_t s_rng[VS_RNG_DEG * VS_NT];
    const uint32_t s = blockIdx.x * VS_NT + threadIdx.x;
    if (s >= a.n_streams) return;
    const VsStream st = a.streams[s];
    VsRng g;
    g.r = s_rng + threadIdx.x;
    vs_rng_seed(g, st.seed);                                             /* flowgen_shimmer.c:241 */

    const bool do_jit = (st.flags & VS_F_JITTER) && st.jitter != 0.0f;    /* :248 */
    const bool do_shm = (st.flags & VS_F_SHIMMER) && st.shimmer != 0.0f;  /* :295 */
    const bool noise = (st.flags & VS_F_NOISE) != 0;                      /* :373 */
    const bool pulse = LOG || noise;
    const int P = st.P, T2 = st.T2;
    const float Pf = (float)P, ampf = (float)st.amp;
    const float t_hi = __fmul_rn(1.2f, Pf), t_lo = __fmul_rn(0.8f, Pf);
    const float a_hi = __fmul_rn(1.8f, ampf), a_lo = __fmul_rn(0.2f, ampf);
    const double jit = (double)st.jitter, shm = (double)st.shimmer;
    const double *ct = a.costab + st.cos_off;
    const float DC = st.DC;
    const int16_t DCs = st.DCs;

    int T = P, T4 = 0, ndw = 0;
    float dper = 0.0f, dsh = 0.0f;
    uint32_t count = 0, np = 0, next_c = 0;
    VsPeriod *tab = a.table + st.tab_off;
    VsChunk *chunks = a.chunks + st.chunk0;
    vs_period_rec *log = LOG ? (vs_period_rec *)a.log + st.log_off : nullptr;
    int guard = 0;

    do {
        uint32_t nd = 0;
        if (do_jit) {                                                     /* :276-290 */
            const float prev = dper;
            float cur;
            do {
                const int32_t r = vs_rng_next(g); nd++;
                double t = __ddiv_rn((double)r, VS_RAND_MAX_D * 10000.0);
                t = __dmul_rn(__dmul_rn(t, 40000.0), jit);
                const float J = __double2float_rn(__dsub_rn(t, __dmul_rn(2.0, jit)));
                const double den = __dsub_rn(2.0, (double)J);
                const double q1 = __ddiv_rn(__dmul_rn((double)prev, __dadd_rn(2.0, (double)J)), den);
                const double q2 = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, (double)P), (double)J), den);
                cur = __double2float_rn(__dadd_rn(q1, q2));
                T = (int)vs_d2s(ceil((double)__fadd_rn(Pf, cur)));
                if (++guard > (1 << 22)) { atomicExch(a.status, VS_ERANGE); return; }
            } while ((float)T > t_hi || (float)T < t_lo);
            dper = cur;
        }
        float A = ampf, S = 0.0f;
        if (do_shm) {                                                     /* :296-306 */
            const float prev = dsh;
            float cur;
            do {
                const int32_t r = vs_rng_next(g); nd++;
                const float eps = __fdiv_rn((float)r, 2147483648.0f);     /* (float)RAND_MAX == 2^31 */
                S = __double2float_rn(__dsub_rn(__dmul_rn(__dmul_rn((double)eps, 4.0), shm), __dmul_rn(2.0, shm)));
                const double den = __dsub_rn(2.0, (double)S);
                const double q1 = __ddiv_rn(__dmul_rn((double)prev, __dadd_rn(2.0, (double)S)), den);
                const double q2 = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, (double)st.amp), (double)S), den);
                cur = __double2float_rn(__dadd_rn(q1, q2));
                A = __fadd_rn(ampf, cur);
                if (++guard > (1 << 22)) { atomicExch(a.status, VS_ERANGE); return; }
            } while (A > a_hi || A < a_lo);
            dsh = cur;
        }
        if (T < 1) { atomicExch(a.status, VS_ERANGE); return; }

        /* closure-speed draw, always consumed (:325) */
        const int32_t rk = vs_rng_next(g); nd++;
        const double kq = __dsub_rn(__ddiv_rn((double)rk, VS_RAND_MAX_D), 0.5);
        const float Knew = __double2float_rn(
            __dmul_rn((double)st.K, __dadd_rn(1.0, __dmul_rn((double)__fmul_rn(2.0f, st.Kvar), kq))));

        /* chunks whose generation starts inside this period: remember where we are */
        while (next_c < st.n_chunks && chunks[next_c].gen_target < count + (uint32_t)T) {
            chunks[next_c].first_period = np;
            if (noise && a.rng_snap) vs_rng_save(g, a.rng_snap + (size_t)(st.chunk0 + next_c) * 32);
            next_c++;
        }

        int T3 = 2 * T2;
        float x_pow = 0.0f, w_pow = 0.0f;
        if (pulse) {
            /* one pass over the open phase: T4 = last rising index below DC (:320-323), T3 = first
             * falling index below DC (:329), and the float power sum over [T4,T3) in index order
             * (:374-378).  The sum restarts whenever T4 moves; if T4 never moves in this period the
             * sum that started at the stale T4 is the one the reference computes. */
            const double Ad = (double)A, Ah = __dmul_rn(Ad, 0.5), Kd = (double)Knew;
            float aux_new = 0.0f, aux_old = 0.0f;
            bool moved = false;
            const int T4_old = T4;
            for (int i = 0; i < T2; i++) {
                int16_t x = vs_rising(Ah, __ldg(ct + i));
                if ((float)x < DC) { x = DCs; T4 = i; moved = true; aux_new = 0.0f; }
                const float sq = __fmul_rn((float)x, (float)x);
                if (moved) aux_new = __fadd_rn(aux_new, sq);
                else if (i >= T4_old) aux_old = __fadd_rn(aux_old, sq);
            }
            float aux = moved ? aux_new : aux_old;
            int i;
            for (i = T2; i < 2 * T2; i++) {
                const int16_t x = vs_falling(Ad, Kd, __ldg(ct + i - T2));
                if ((float)x < DC) break;
                aux = __fadd_rn(aux, __fmul_rn((float)x, (float)x));
            }
            T3 = i;
            if (noise) {                                                  /* :378-382 */
                const float span = __fsub_rn((float)T3, (float)T4);
                x_pow = __fdiv_rn(aux, span);
                const float ax = __double2float_rn(__dadd_rn(1.0, (double)__fdiv_rn(span, (float)T)));
                ndw = vs_d2i(sqrt((double)__fdiv_rn(__fmul_rn(__fmul_rn(12.0f, ax), x_pow), st.noise)));
                const int n1 = T4, n2 = T > T3 ? T - T3 : 0;
                if (LOG) {
                    float wa = 0.0f;
                    for (int k = 0; k < n1 + n2; k++) {
                        const int16_t w = vs_noise_w(vs_rng_next(g), ndw);
                        wa = __fadd_rn(wa, __fmul_rn((float)w, (float)w));
                    }
                    w_pow = __fdiv_rn(wa, (float)T);
                } else {
                    for (int k = 0; k < n1 + n2; k++) (void)vs_rng_next(g);
                }
                nd += (uint32_t)(n1 + n2);
            }
        }

        if (np >= st.tab_cap) { atomicExch(a.status, VS_ENOMEM); return; }
        VsPeriod e;
        e.start = count; e.T = T; e.A = A; e.Knew = Knew; e.T3 = T3; e.T4 = T4; e.ndw = ndw;
        e.npert = noise ? nd - (uint32_t)(T4 + (T > T3 ? T - T3 : 0)) : nd;
        tab[np] = e;
        if (LOG) {
            vs_period_rec r;
            r.T = T; r.T2 = T2; r.T3 = T3; r.T4 = T4; r.A = A; r.Knew = Knew; r.S = S;
            r.ndraws = (int32_t)nd; r.ndw = ndw; r.x_pow = x_pow; r.w_pow = w_pow; r.reserved = 0;
            r.start = count;
            log[np] = r;
        }
        count += (uint32_t)T;                                             /* :413 */
        np++;
    } while (count < st.n);                                               /* :423 */
    a.n_periods[s] = np;
}

/* ================================================================================================
 * RENDER: one thread per (stream, chunk)
 * ============================================================================================== */
enum { VS_MODE_FLOW = 0, VS_MODE_SYNTH = 1, VS_MODE_FILTER = 2 };

template <bool NOISE>
struct VsFlowGen {
    const VsPeriod *tab;
    const double *ct;
    uint32_t pidx;
    int i, T, T2, T3, T4, ndw;
    bool closed, noise;
    double Ah, Ad, Kd;
    float DC;
    int16_t DCs;
    VsRng g;

    __device__ __forceinline__ void load_period(bool skip_pert)
    {
        const int4 *q = reinterpret_cast<const int4 *>(tab + pidx);
        const int4 lo = __ldg(q), hi = __ldg(q + 1);
        T = lo.y;
        Ad = (double)__int_as_float(lo.z);
        Ah = __dmul_rn(Ad, 0.5);
        Kd = (double)__int_as_float(lo.w);
        T3 = hi.x; T4 = hi.y; ndw = hi.z;
        i = 0; closed = false;
        if (NOISE && noise && skip_pert)
            for (uint32_t k = 0; k < (uint32_t)hi.w; k++) (void)vs_rng_next(g);
    }

    __device__ __forceinline__ int16_t next()
    {
        while (i >= T) { pidx++; load_period(true); }
        int16_t x;
        if (i < T2) {                                                     /* :318-324 */
            x = vs_rising(Ah, __ldg(ct + i));
            if ((float)x < DC) x = DCs;
        } else if (!closed && i < 2 * T2) {                               /* :327-330 */
            x = vs_falling(Ad, Kd, __ldg(ct + i - T2));
            if ((float)x < DC) { closed = true; x = DCs; }
        } else {
            x = DCs;                                                      /* :334-336 */
        }
        if (NOISE && noise && (i < T4 || i >= T3)) {                      /* :385-406 */
            const int16_t w = vs_noise_w(vs_rng_next(g), ndw);
            x = vs_clip_ceil(__fadd_rn((float)x, (float)w));
        }
        i++;
        return x;
    }
};

/* vowel_new.c:413-427, literally */
__device__ __forceinline__ int16_t vs_round2int(double v)
{
    const double dec = __dsub_rn(v, floor(v));
    if (dec > 0.5) v = __dadd_rn(v, 1.0);
    if (v > 32767.0) v = 32767.0;
    else if (v < -32767.0) v = -32767.0;
    return vs_d2s(floor(v));
}

__device__ __forceinline__ uint32_t vs_pack2(int16_t lo, int16_t hi)
{
    return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16);
}

template <int MODE, bool NOISE, bool EXACT>
__global__ void __launch_bounds__(VS_NT) vs_render_kernel(const VsRenderArgs a)
{
    __shared__ int32_t s_x[VS_RING * VS_NT];
    __shared__ uint32_t s_rng[(NOISE && MODE != VS_MODE_FILTER) ? VS_RNG_DEG * VS_NT : 1];

    const uint32_t c = blockIdx.x * VS_NT + threadIdx.x;
    if (c >= a.n_chunks) return;
    const VsChunk ck = a.chunks[c];
    const VsStream st = a.streams[ck.stream];
    int32_t *xs = s_x + threadIdx.x;

    /* ---- flow source ---- */
    VsFlowGen<NOISE> gen;
    int64_t nstart;                       /* first sample this thread generates / reads */
    const int16_t *fin = nullptr;
    if (MODE == VS_MODE_FILTER) {
        nstart = ck.gen_target;
        fin = a.flow_in + st.in_off;
    } else {
        gen.tab = a.table + st.tab_off;
        gen.ct = a.costab + st.cos_off;
        gen.pidx = ck.first_period;
        gen.T2 = st.T2; gen.DC = st.DC; gen.DCs = st.DCs;
        gen.noise = (st.flags & VS_F_NOISE) != 0;
        gen.g.r = s_rng + (NOISE ? threadIdx.x : 0);
        gen.g.f = 3;
        if (NOISE && gen.noise) vs_rng_load(gen.g, a.rng_snap + (size_t)c * 32);
        gen.load_period(false);           /* the snapshot was taken after this period's K draw */
        nstart = __ldg(&gen.tab[gen.pidx].start);
    }

    /* ---- filter state: 24-entry ring in registers, coefficients in registers ---- */
    double y[VS_RING];
    double cf[VS_ORDER + 1];
    double gaind = 0.0, pred = 0.0;
    if (MODE != VS_MODE_FLOW) {
#pragma unroll
        for (int j = 0; j < VS_RING; j++) y[j] = 0.0;
#pragma unroll
        for (int j = 0; j <= VS_ORDER; j++) cf[j] = __ldg(a.coef + st.preset * VS_RING + j);
        gaind = (double)st.gain;
        pred = (double)st.pre;
    }

    /* ---- output geometry: blocks of 24 samples anchored on 16-byte boundaries of the row ---- */
    int16_t *orow = a.pcm_out + st.out_off;
    double *rrow = a.raw_out ? a.raw_out + st.out_off : nullptr;
    const int64_t phase = (int64_t)((reinterpret_cast<uintptr_t>(orow) >> 1) & 7);
    const int64_t lo = ck.emit_lo, hi = ck.emit_hi;
    int64_t blk = nstart - ((phase + nstart) & 7);

    for (; blk < hi; blk += VS_RING) {
        /* 1. 24 flow samples -> shared memory (rolled: the generator is a divergent state machine) */
#pragma unroll 1
        for (int k = 0; k < VS_RING; k++) {
            const int64_t m = blk + k;
            int32_t x = 0;
            if (m >= nstart && m < hi) x = (MODE == VS_MODE_FILTER) ? (int32_t)__ldg(fin + m) : (int32_t)gen.next();
            xs[k * VS_NT] = x;
        }

        /* 2. recurrence, fully unrolled so that the ring indices are compile-time registers */
        int16_t q[VS_RING];
#pragma unroll
        for (int k = 0; k < VS_RING; k++) {
            const int32_t xi = xs[k * VS_NT];
            if (MODE == VS_MODE_FLOW) {
                q[k] = (int16_t)xi;
            } else {
                double acc = __dmul_rn((double)xi, gaind);                /* vowel_new.c:266-269 */
                double v;
                if (EXACT) {
#pragma unroll
                    for (int j = 1; j <= VS_ORDER; j++)                   /* :279-281, same order */
                        acc = __dsub_rn(acc, __dmul_rn(cf[j], y[(k + VS_RING - j) % VS_RING]));
                    v = __dsub_rn(acc, __dmul_rn(pred, y[(k + VS_RING - 1) % VS_RING]));   /* :284 */
                } else {
#pragma unroll
                    for (int j = VS_ORDER; j >= 1; j--)                   /* oldest tap first */
                        acc = fma(-cf[j], y[(k + VS_RING - j) % VS_RING], acc);
                    v = fma(-pred, y[(k + VS_RING - 1) % VS_RING], acc);
                }
                y[k] = acc;                                               /* :287-289 (ring) */
                q[k] = vs_round2int(v);
                if (rrow) {
                    const int64_t m = blk + k;
                    if (m >= lo && m < hi) rrow[m] = v;
                }
            }
        }

        /* 3. three 16-byte pieces; interior pieces are whole by construction */
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int64_t m0 = blk + 8 * p;
            if (m0 >= lo && m0 + 8 <= hi) {
                uint4 w;
                w.x = vs_pack2(q[8 * p + 0], q[8 * p + 1]);
                w.y = vs_pack2(q[8 * p + 2], q[8 * p + 3]);
                w.z = vs_pack2(q[8 * p + 4], q[8 * p + 5]);
                w.w = vs_pack2(q[8 * p + 6], q[8 * p + 7]);
                *reinterpret_cast<uint4 *>(orow + m0) = w;
            } else if (m0 + 8 > lo && m0 < hi) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int64_t m = m0 + k;
                    if (m >= lo && m < hi) orow[m] = q[8 * p + k];
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * launch wrappers (called from vs_api.cu)
 * ---------------------------------------------------------------------------------------------- */
cudaError_t vs_launch_plan(const VsPlanArgs &a, bool want_log, cudaStream_t s)
{
    const unsigned grid = (a.n_streams + VS_NT - 1) / VS_NT;
    if (want_log) vs_plan_kernel<true><<<grid, VS_NT, 0, s>>>(a);
    else vs_plan_kernel<false><<<grid, VS_NT, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t vs_launch_render(const VsRenderArgs &a, int mode, bool noise, bool exact, cudaStream_t s)
{
    const unsigned grid = (a.n_chunks + VS_NT - 1) / VS_NT;
#define VS_GO(M, N, E) vs_render_kernel<M, N, E><<<grid, VS_NT, 0, s>>>(a)
    if (mode == VS_MODE_FLOW) { if (noise) VS_GO(VS_MODE_FLOW, true, false); else VS_GO(VS_MODE_FLOW, false, false); }
    else if (mode == VS_MODE_FILTER) { if (exact) VS_GO(VS_MODE_FILTER, false, true); else VS_GO(VS_MODE_FILTER, false, false); }
    else {
        if (noise) { if (exact) VS_GO(VS_MODE_SYNTH, true, true); else VS_GO(VS_MODE_SYNTH, true, false); }
        else       { if (exact) VS_GO(VS_MODE_SYNTH, false, true); else VS_GO(VS_MODE_SYNTH, false, false); }
    }
#undef VS_GO
    return cudaGetLastError();
}
