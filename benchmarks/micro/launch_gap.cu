// What spaces back-to-back small kernels on this box, and what removes it?  64 MiB cold reads (the C1 sum_axis input):
//   plain    12 launches in a stream
//   graph    the same 12 launches captured once and replayed as a CUDA graph
//   pdl      programmatic dependent launch: every kernel starts with griddepcontrol.launch_dependents + griddepcontrol.wait,
//            launched with cudaLaunchAttributeProgrammaticStreamSerialization (its blocks get scheduled while the previous
//            kernel drains and wait for its completion before touching memory)
//   pdl+graph
// and the bare launch rate of an empty kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_gap launch_gap.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void __launch_bounds__(256) rd(const uint4 *__restrict__ p, size_t nvec, float *out)
{
    if (PDL) {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p[i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].y) + __uint_as_float(v[u].z) + __uint_as_float(v[u].w);
    }
    for (; i < nvec; i += stride) {
        uint4 v = p[i];
        acc += __uint_as_float(v.x) + __uint_as_float(v.y) + __uint_as_float(v.z) + __uint_as_float(v.w);
    }
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + blockIdx.x, acc);
}

__global__ void empty_kernel() {}

static cudaStream_t st;

template <bool PDL>
static void launch(const void *buf, size_t nvec, float *out, int blocks)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = PDL ? 1 : 0;
    cudaLaunchKernelEx(&cfg, rd<PDL>, (const uint4 *)buf, nvec, out);
}

int main(int argc, char **argv)
{
    const size_t bytes = (argc > 1 ? atol(argv[1]) : 64) << 20;
    const size_t nvec = bytes / 16;
    const int blocks = argc > 2 ? atoi(argv[2]) : 1184;
    void *buf[4];
    float *out;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (auto &b : buf) {
        cudaMalloc(&b, bytes);
        cudaMemset(b, 0, bytes);
    }
    cudaMalloc(&out, 1 << 20);
    cudaMemset(out, 0, 1 << 20);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto timeit = [&](const char *label, auto &&body, int per) {
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
            t.push_back(ms / per * 1e3f);
        }
        std::sort(t.begin(), t.end());
        printf("%-28s %7.2f us per kernel (best %.2f)  %6.0f GB/s\n", label, t[4], t[0], bytes / (t[4] * 1e-6) / 1e9);
    };
    printf("bytes %zu blocks %d\n", bytes, blocks);
    timeit("empty kernel x100", [&] { for (int k = 0; k < 100; ++k) empty_kernel<<<1, 32, 0, st>>>(); }, 100);
    timeit("plain x12", [&] { for (int k = 0; k < 12; ++k) launch<false>(buf[k & 3], nvec, out, blocks); }, 12);
    timeit("pdl x12", [&] { for (int k = 0; k < 12; ++k) launch<true>(buf[k & 3], nvec, out, blocks); }, 12);
    for (int pdl = 0; pdl < 2; ++pdl) {
        cudaGraph_t g;
        cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < 12; ++k) {
            if (pdl) launch<true>(buf[k & 3], nvec, out, blocks);
            else launch<false>(buf[k & 3], nvec, out, blocks);
        }
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        timeit(pdl ? "graph of 12, pdl edges" : "graph of 12", [&] { cudaGraphLaunch(ge, st); }, 12);
        cudaGraphExecDestroy(ge);
        cudaGraphDestroy(g);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
