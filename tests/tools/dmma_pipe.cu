// Is the FP64 tensor path (mma.sync m8n8k4 / m16n8k8 f64) a second FP64 pipe next to DFMA on B200, or the same units?
// Per SM sub-partition: one warp of DFMA chains, one warp of DMMA chains, then both together.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// role per warp: 0 = DFMA (8 independent chains x 3), 1 = DMMA m8n8k4 (8 independent accumulators), 2 = DMMA m16n8k8 (4 accumulators)
__global__ void k(double *out, int iters, const double *g, long long *cyc, int role_even, int role_odd, int warps_active)
{
    const int warp = threadIdx.x >> 5;
    if (warp >= warps_active) return;
    const int role = (warp & 4) ? role_odd : role_even;     // warps 0-3: one per SMSP (even role), warps 4-7: second per SMSP
    double x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = g[(threadIdx.x + i) & 255]; y[i] = g[(threadIdx.x + 3 * i + 7) & 255]; }
    const double c0 = g[5], c1 = g[9];
    long long t0 = clock64();
    if (role == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = __fma_rn(y[i], c0, x[i]);
        }
    } else if (role == 1) {
        double d[16];
#pragma unroll
        for (int i = 0; i < 16; i++) d[i] = x[i & 7];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) dmma884(d[2 * i], d[2 * i + 1], y[i], c1);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) x[i & 7] += d[i];
    } else {
        double d[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) d[i][j] = x[(i + j) & 7];
        double a[4] = {y[0], y[1], y[2], y[3]}, b[2] = {y[4], y[5]};
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 4; i++) dmma1688(d[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) x[(i + j) & 7] += d[i][j];
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 123.456) out[0] = s;
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
}

static void run(const char *name, const double *g, int re, int ro, int warps)
{
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 64);
    const int iters = 4096;
    for (int r = 0; r < 2; r++) k<<<148, 256>>>(out, iters, g, cyc, re, ro, warps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    auto per = [&](int role, long long c) { return (double)c / iters / (role == 0 ? 24 : role == 1 ? 8 : 4); };
    auto flop = [&](int role) { return role == 0 ? 64.0 : role == 1 ? 512.0 : 2048.0; };
    printf("%-58s warp0: %.2f cyc/instr (%.1f flop/clk/SMSP)", name, per(re, h[0]), flop(re) / per(re, h[0]));
    if (warps > 4) printf("   warp4: %.2f cyc/instr (%.1f flop/clk/SMSP)", per(ro, h[4]), flop(ro) / per(ro, h[4]));
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    double h[256]; for (int i = 0; i < 256; i++) h[i] = 0.999 + 1e-6 * i;
    double *g; cudaMalloc(&g, sizeof h); cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
    run("DFMA alone (1 warp/SMSP)", g, 0, 0, 4);
    run("DMMA m8n8k4 alone (1 warp/SMSP)", g, 1, 1, 4);
    run("DMMA m16n8k8 alone (1 warp/SMSP)", g, 2, 2, 4);
    run("DFMA + DFMA (2 warps/SMSP)", g, 0, 0, 8);
    run("DMMA884 + DMMA884 (2 warps/SMSP)", g, 1, 1, 8);
    run("DFMA + DMMA m8n8k4 (2 warps/SMSP)", g, 0, 1, 8);
    run("DFMA + DMMA m16n8k8 (2 warps/SMSP)", g, 0, 2, 8);
    return 0;
}
