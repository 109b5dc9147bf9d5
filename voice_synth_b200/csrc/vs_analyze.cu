/* vs_analyze.cu -- SURVEY.md 8f row N4: cycle-to-cycle analysis of generated glottal flow on the GPU (F0, local
 * jitter, local shimmer), the measurements the reference's missing `acoustic` tools are described to make
 * (/root/reference README:14-16; no code in the tree, so the definitions are this library's own and are stated in
 * include/voicesynth.h; tests/analysis_ref.py restates them in numpy).
 *
 *   vs_onsets_kernel   one warp per stream, lanes along the stream, 32 samples per step.  A two-threshold trigger
 *                      (armed by x <= lo, fired by x > hi) is a scan over the composition of per-sample state maps
 *                      {arm, disarm, keep}: the state in front of a sample is that of the LAST arm/disarm before it
 *                      -- two ballots and one unsigned compare of the masked masks.  Onsets are written compactly.
 *   vs_cycles_kernel   one warp per stream, lanes along the cycles: length and peak of each cycle, exact integer
 *                      sums of lengths, peaks and absolute first differences, statistics in FP64 by lane 0.
 */
#include "vs_device.cuh"

#define VS_AN_NT 128

__global__ void __launch_bounds__(VS_AN_NT) vs_onsets_kernel(const int16_t *flow, const VsAnalyzeRow *rows, uint32_t n_rows,
                                                             uint32_t *onsets, uint32_t *counts)
{
    const int lane = threadIdx.x & 31;
    const uint32_t s = blockIdx.x * (VS_AN_NT / 32) + (threadIdx.x >> 5);
    if (s >= n_rows) return;
    const VsAnalyzeRow r = rows[s];
    const int16_t *x = flow + r.off;
    uint32_t *ons = onsets + r.ons_off;
    const uint32_t below = (1u << lane) - 1u;
    bool armed = true;                                      /* a stream begins in its closed phase */
    uint32_t count = 0;
    for (uint32_t base = 0; base < r.n; base += 32) {
        const uint32_t m = base + (uint32_t)lane;
        const bool valid = m < r.n;
        const int v = valid ? (int)x[m] : 0;
        const uint32_t A = __ballot_sync(VS_FULL, valid && v <= (int)r.lo);     /* arms */
        const uint32_t D = __ballot_sync(VS_FULL, valid && v > (int)r.hi);      /* fires if armed, disarms */
        const uint32_t la = A & below, ld = D & below;
        /* the masks are disjoint: the one with the higher top bit is the larger number */
        const bool armed_here = (la | ld) ? la > ld : armed;
        const bool fire = ((D >> lane) & 1u) && armed_here;
        const uint32_t F = __ballot_sync(VS_FULL, fire);
        if (fire) {
            const uint32_t k = count + (uint32_t)__popc(F & below);
            if (k < r.cap) ons[k] = m;
        }
        count += (uint32_t)__popc(F);
        if (A | D) armed = A > D;
    }
    if (lane == 0) counts[s] = count;
}

__device__ __forceinline__ long long vs_warp_sum(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(VS_FULL, v, o);
    return v;
}

__global__ void __launch_bounds__(VS_AN_NT) vs_cycles_kernel(const int16_t *flow, const VsAnalyzeRow *rows, uint32_t n_rows,
                                                             const uint32_t *onsets, const uint32_t *counts, vs_flow_stats *out)
{
    const int lane = threadIdx.x & 31;
    const uint32_t s = blockIdx.x * (VS_AN_NT / 32) + (threadIdx.x >> 5);
    if (s >= n_rows) return;
    const VsAnalyzeRow r = rows[s];
    const int16_t *x = flow + r.off;
    const uint32_t *ons = onsets + r.ons_off;
    const uint32_t K = counts[s];
    vs_flow_stats st;
    st.cycles = 0; st.flags = 0; st.f0_hz = 0.0f; st.jitter_pct = 0.0f; st.shimmer_pct = 0.0f; st.mean_period = 0.0f; st.mean_peak = 0.0f;
    st.onsets = K;
    if (K > r.cap) {                                        /* more onsets than the scratch holds: thresholds inside the noise? */
        st.flags = VS_STATS_OVERFLOW;
        if (lane == 0) out[s] = st;
        return;
    }
    const uint32_t nc = K >= 2 ? K - 1 : 0;                 /* complete cycles */
    long long sT = 0, sP = 0, sdT = 0, sdP = 0;
    int carry_peak = 0;                                     /* peak of the cycle before this step's first */
    for (uint32_t k0 = 0; k0 < nc; k0 += 32) {
        const uint32_t k = k0 + (uint32_t)lane;
        const bool valid = k < nc;
        int T = 0, peak = -32768, Tprev = 0;
        if (valid) {
            const uint32_t o0 = ons[k], o1 = ons[k + 1];
            T = (int)(o1 - o0);
            if (k > 0) Tprev = (int)(o0 - ons[k - 1]);
            for (uint32_t m = o0; m < o1; m++) peak = max(peak, (int)x[m]);
        }
        int pprev = __shfl_up_sync(VS_FULL, peak, 1);
        if (lane == 0) pprev = carry_peak;
        carry_peak = __shfl_sync(VS_FULL, peak, 31);
        if (valid) {
            sT += T; sP += peak;
            if (k > 0) { sdT += abs(T - Tprev); sdP += abs(peak - pprev); }
        }
    }
    sT = vs_warp_sum(sT); sP = vs_warp_sum(sP); sdT = vs_warp_sum(sdT); sdP = vs_warp_sum(sdP);
    if (lane == 0) {
        st.cycles = nc;
        if (nc >= 1) {
            const double mT = (double)sT / (double)nc, mP = (double)sP / (double)nc;
            st.mean_period = (float)mT;
            st.mean_peak = (float)mP;
            st.f0_hz = (float)((double)r.fs / mT);
            if (nc >= 2) {
                st.jitter_pct = (float)(100.0 * ((double)sdT / (double)(nc - 1)) / mT);
                st.shimmer_pct = mP != 0.0 ? (float)(100.0 * ((double)sdP / (double)(nc - 1)) / mP) : 0.0f;
            }
        }
        out[s] = st;
    }
}

cudaError_t vs_launch_analyze(const int16_t *flow, const VsAnalyzeRow *rows, uint32_t n_rows, uint32_t *onsets, uint32_t *counts,
                              vs_flow_stats *out, cudaStream_t s)
{
    const unsigned grid = (n_rows + VS_AN_NT / 32 - 1) / (VS_AN_NT / 32);
    vs_onsets_kernel<<<grid, VS_AN_NT, 0, s>>>(flow, rows, n_rows, onsets, counts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    vs_cycles_kernel<<<grid, VS_AN_NT, 0, s>>>(flow, rows, n_rows, onsets, counts, out);
    return cudaGetLastError();
}
