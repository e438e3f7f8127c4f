// Row-shaped reads of a cold f32 [4096, 4096] array (64 MiB): one warp (or a pair of rows per warp) per 16 KiB row, the shape of
// sum_axis(Axis(1)) in config 1.  Variants: warps per block, vectors in flight per lane, rows handled per warp in parallel.
// All launched with PDL, 12 back to back over 4 rotating buffers; the per-row result goes to out[row].
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int U, int RPW>
__global__ void __launch_bounds__(256) rd_rows(const uint4 *__restrict__ p, int row_vec, int nrows, float *out)
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int row0 = (blockIdx.x * wpb + warp) * RPW;
    if (row0 >= nrows) return;
    float acc[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) acc[r] = 0.f;
    for (int i = lane; i < row_vec; i += 32 * U) {
        uint4 v[RPW][U];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int u = 0; u < U; ++u) v[r][u] = p[(size_t)(row0 + r) * row_vec + i + 32 * u];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int u = 0; u < U; ++u)
                acc[r] += __uint_as_float(v[r][u].x) + __uint_as_float(v[r][u].y) + __uint_as_float(v[r][u].z) + __uint_as_float(v[r][u].w);
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        float a = acc[r];
        for (int d = 16; d > 0; d >>= 1) a += __shfl_down_sync(0xffffffffu, a, d);
        if (lane == 0) out[row0 + r] = a;
    }
}

template <int U, int RPW>
static float run(int wpb, void **buf, float *out, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1)
{
    const int n = 4096, row_vec = n / 4;
    const int blocks = (n / RPW + wpb - 1) / wpb;
    auto body = [&] {
        for (int k = 0; k < 12; ++k) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(blocks);
            cfg.blockDim = dim3(32 * wpb);
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, rd_rows<U, RPW>, (const uint4 *)buf[k & 3], row_vec, n, out);
        }
    };
    body();
    cudaStreamSynchronize(st);
    std::vector<float> t;
    for (int rep = 0; rep < 9; ++rep) {
        cudaEventRecord(e0, st);
        body();
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        t.push_back(ms / 12 * 1e3f);
    }
    std::sort(t.begin(), t.end());
    return t[4];
}

int main()
{
    const size_t bytes = (size_t)4096 * 4096 * 4;
    void *buf[4];
    float *out;
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (auto &b : buf) {
        cudaMalloc(&b, bytes);
        cudaMemset(b, 0, bytes);
    }
    cudaMalloc(&out, 1 << 20);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("U rows/warp warps/block us GB/s\n");
    for (int wpb : {2, 4, 8}) {
        float t;
        t = run<4, 1>(wpb, buf, out, st, e0, e1); printf("4 1 %d %6.2f %5.0f\n", wpb, t, bytes / (t * 1e-6) / 1e9);
        t = run<8, 1>(wpb, buf, out, st, e0, e1); printf("8 1 %d %6.2f %5.0f\n", wpb, t, bytes / (t * 1e-6) / 1e9);
        t = run<16, 1>(wpb, buf, out, st, e0, e1); printf("16 1 %d %6.2f %5.0f\n", wpb, t, bytes / (t * 1e-6) / 1e9);
        t = run<4, 2>(wpb, buf, out, st, e0, e1); printf("4 2 %d %6.2f %5.0f\n", wpb, t, bytes / (t * 1e-6) / 1e9);
        t = run<8, 2>(wpb, buf, out, st, e0, e1); printf("8 2 %d %6.2f %5.0f\n", wpb, t, bytes / (t * 1e-6) / 1e9);
        t = run<2, 4>(wpb, buf, out, st, e0, e1); printf("2 4 %d %6.2f %5.0f\n", wpb, t, bytes / (t * 1e-6) / 1e9);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
