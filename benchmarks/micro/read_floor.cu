// How fast can ANY kernel read a 64 MiB array (BASELINE config 1's sum_axis input) that is cold in L2?  The floor the C1
// reductions are judged against: grid-stride 128-bit loads, U independent loads in flight per thread, one float per block out.
// Four 64 MiB buffers are rotated (256 MiB > the 126 MB L2) and 12 launches go back to back between one event pair.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_floor read_floor.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(512) rd(const uint4 *__restrict__ p, size_t nvec, float *out)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = p[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].y) + __uint_as_float(v[u].z) + __uint_as_float(v[u].w);
    }
    for (; i < nvec; i += stride) {
        uint4 v = p[i];
        acc += __uint_as_float(v.x) + __uint_as_float(v.y) + __uint_as_float(v.z) + __uint_as_float(v.w);
    }
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    __shared__ float s[16];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s[w];
        out[blockIdx.x] = t;
    }
}

int main(int argc, char **argv)
{
    const size_t bytes = (argc > 1 ? atol(argv[1]) : 64) << 20;
    const size_t nvec = bytes / 16;
    void *buf[4];
    float *out;
    for (auto &b : buf) {
        cudaMalloc(&b, bytes);
        cudaMemset(b, 0, bytes);
    }
    cudaMalloc(&out, 1 << 20);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("bytes %zu\nblocks threads U us_per_launch GB/s\n", bytes);
    for (int threads : {256, 512})
        for (int blocks : {148, 296, 444, 592, 888, 1184, 2368, 4736, 9472}) {
            for (int U : {4, 8}) {
                auto launch = [&](int k) {
                    const uint4 *p = (const uint4 *)buf[k & 3];
                    if (U == 4) rd<4><<<blocks, threads>>>(p, nvec, out);
                    else rd<8><<<blocks, threads>>>(p, nvec, out);
                };
                for (int k = 0; k < 8; ++k) launch(k);
                std::vector<float> t;
                for (int rep = 0; rep < 7; ++rep) {
                    cudaEventRecord(e0);
                    for (int k = 0; k < 12; ++k) launch(k);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    t.push_back(ms / 12 * 1e3f);
                }
                std::sort(t.begin(), t.end());
                printf("%5d %4d %d %7.2f %7.0f\n", blocks, threads, U, t[3], bytes / (t[3] * 1e-6) / 1e9);
            }
        }
    return 0;
}
