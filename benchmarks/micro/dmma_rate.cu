// DMMA.8x8x4 issue-rate microbenchmark: cycles per DMMA per SM sub-partition as a function of warps per sub-partition,
// number of independent accumulators and operand pattern.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MI x NJ accumulator tiles, one sweep = MI*NJ DMMAs; PAT 0: i outer / j inner, 1: j outer / i inner
template <int MI, int NJ, int PAT>
__global__ void k(double *out, const double *in, int iters, long long *cycles)
{
    double acc[MI][NJ][2];
    double a[MI], b[NJ];
    for (int i = 0; i < MI; ++i) a[i] = in[threadIdx.x + 32 * i];
    for (int j = 0; j < NJ; ++j) b[j] = in[threadIdx.x + 32 * (MI + j)];
    for (int i = 0; i < MI; ++i)
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (PAT == 0) {
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        } else {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int i = 0; i < MI; ++i) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < MI; ++i)
        for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MI, int NJ, int PAT>
void run(const char *label, double *out, double *in, long long *cyc)
{
    const int iters = getenv("ITERS") ? atoi(getenv("ITERS")) : 2000;
    for (int warps = 4; warps <= 8; warps += 4) {
        k<MI, NJ, PAT><<<148, warps * 32>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<MI, NJ, PAT><<<148, warps * 32>>>(out, in, iters, cyc);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double per_smsp = (double)iters * MI * NJ * (warps / 4);
        printf("%-22s warps/SMSP %d: %6.2f clk per DMMA per SMSP (clock64), %6.1f TFLOP/s (events)\n", label, warps / 4,
               c / per_smsp, 148.0 * warps * iters * MI * NJ * 512.0 / (ms * 1e-3) / 1e12);
    }
}

int main()
{
    double *out, *in;
    long long *cyc;
    cudaMalloc(&out, 148 * 512 * 8);
    cudaMalloc(&in, 4096 * 8);
    cudaMemset(in, 0, 4096 * 8);
    cudaMalloc(&cyc, 8);
    run<8, 4, 0>("8x4 i-outer", out, in, cyc);
    run<8, 4, 1>("8x4 j-outer", out, in, cyc);
    run<4, 8, 0>("4x8 i-outer", out, in, cyc);
    run<4, 4, 0>("4x4 i-outer", out, in, cyc);
    run<2, 2, 0>("2x2", out, in, cyc);
    run<1, 1, 0>("1x1 (dependent chain)", out, in, cyc);
    return 0;
}
