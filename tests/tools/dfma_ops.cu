// DFMA issue rate vs number of distinct register operands (1 warp per SM sub-partition, 8 chains).
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cc[64];
template <int MODE>
__global__ void k(double *out, int iters, const double *g, long long *cyc)
{
    double x[8], c[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; c[i] = g[threadIdx.x + i]; d[i] = g[64 + threadIdx.x + i]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) x[i] = __fma_rn(x[i], cc[i], cc[8 + i]);          // 1 register operand
            else if (MODE == 1) x[i] = __fma_rn(x[i], c[i], cc[8 + i]);      // 2 register operands (acc as multiplicand)
            else if (MODE == 2) x[i] = __fma_rn(c[i], cc[i], x[i]);          // 2 register operands (acc as addend), const multiplier
            else if (MODE == 3) x[i] = __fma_rn(c[i], d[i], x[i]);           // 3 register operands, all distinct
            else x[i] = __fma_rn(c[i], d[i & 3], x[i]);                      // 3 register operands, one shared between pairs
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char *name, const double *g)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int r = 0; r < 2; r++) k<MODE><<<148, 128>>>(out, iters, g, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-50s %.2f cycles per DFMA\n", name, (double)h / iters / 8);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    double h[256]; for (int i = 0; i < 256; i++) h[i] = 0.999 + 1e-6 * i;
    double *g; cudaMalloc(&g, sizeof h); cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(cc, h, 64 * 8);
    run<0>("acc*const+const (1 reg operand)", g);
    run<1>("acc*reg+const (2 reg operands)", g);
    run<2>("reg*const+acc (2 reg operands, const multiplier)", g);
    run<3>("reg*reg+acc (3 distinct reg operands)", g);
    run<4>("reg*reg+acc (3 reg operands, shared multiplicand)", g);
    return 0;
}
