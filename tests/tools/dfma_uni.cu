// Can block-uniform coefficients (constant memory indexed via blockIdx, or kernel params) reach the 2-register DFMA rate?
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cc[10][24];
struct P { double c[24]; };
template <int MODE>
__global__ void k(double *out, int iters, const double *g, long long *cyc, const int *sel, P prm)
{
    double x[8], y[24];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 24; i++) y[i] = g[threadIdx.x + i];
    const double *cp = cc[sel[blockIdx.x]];       // block-uniform preset
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) x[i] = __fma_rn(y[j * 8 + i], cp[j * 8 + i], x[i]);
            else x[i] = __fma_rn(y[j * 8 + i], prm.c[j * 8 + i], x[i]);
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
void run(const char *name, const double *g, const int *sel, P prm)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    for (int r = 0; r < 2; r++) k<MODE><<<148, 128>>>(out, iters, g, cyc, sel, prm);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-50s %.2f cycles per DFMA\n", name, (double)h / iters / 24);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    double h[256]; for (int i = 0; i < 256; i++) h[i] = 0.999 + 1e-6 * i;
    double *g; cudaMalloc(&g, sizeof h); cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(cc, h, 240 * 8);
    int hs[148]; for (int i = 0; i < 148; i++) hs[i] = i % 10;
    int *sel; cudaMalloc(&sel, sizeof hs); cudaMemcpy(sel, hs, sizeof hs, cudaMemcpyHostToDevice);
    P prm; for (int i = 0; i < 24; i++) prm.c[i] = h[i];
    run<0>("coef = __constant__[blockIdx-selected preset][j]", g, sel, prm);
    run<1>("coef = kernel parameter", g, sel, prm);
    return 0;
}
