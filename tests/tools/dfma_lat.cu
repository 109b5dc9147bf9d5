// DFMA latency/throughput microbenchmark: W warps per SM sub-partition, ILP independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b, long long *cyc)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = __fma_rn(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps_per_smsp)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<ILP><<<148, 128 * warps_per_smsp>>>(out, iters, 0.999999, 1e-9, cyc);
    k<ILP><<<148, 128 * warps_per_smsp>>>(out, iters, 0.999999, 1e-9, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SMSP %d ILP %2d: %.2f cycles per DFMA per warp-slot (%.2f per chain step)\n", warps_per_smsp, ILP,
           (double)h / iters / ILP / warps_per_smsp, (double)h / iters);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w = 1; w <= 2; w++) { run<1>(w); run<2>(w); run<4>(w); run<6>(w); run<8>(w); run<12>(w); run<16>(w); }
    return 0;
}
