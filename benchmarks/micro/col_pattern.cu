// Does the SHAPE of a column reduction's reads matter for a 64 MiB cold array?  f32 [4096, 4096] read by blocks of 256
// threads arranged (tile_vec, 256 / tile_vec): a block owns tile_vec 16-byte column vectors (tile_vec * 16 contiguous
// bytes per row) and a range of rows; 8 rows in flight per thread.  tile_vec = 32 is red_cols' shape (512 B pieces at a
// 16 KiB pitch), 1024 would be whole rows.  All launched with PDL, 12 back to back over 4 rotating buffers.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) rd_cols(const uint4 *__restrict__ p, int row_vec, int nrows, int tile_vec, int splits, float *out)
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tiles = row_vec / tile_vec;
    const int tile = blockIdx.x % tiles, split = blockIdx.x / tiles;
    const int tx = threadIdx.x % tile_vec, ty = threadIdx.x / tile_vec, ny = 256 / tile_vec;
    const int chunk = (nrows + splits - 1) / splits;
    const int r0 = split * chunk, r1 = min(nrows, r0 + chunk);
    const uint4 *src = p + (size_t)tile * tile_vec + tx;
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int r = r0 + ty;
    for (; r + 7 * ny < r1; r += 8 * ny) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(r + u * ny) * row_vec];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 += __uint_as_float(v[u].x);
            a1 += __uint_as_float(v[u].y);
            a2 += __uint_as_float(v[u].z);
            a3 += __uint_as_float(v[u].w);
        }
    }
    for (; r < r1; r += ny) {
        uint4 v = src[(size_t)r * row_vec];
        a0 += __uint_as_float(v.x);
        a1 += __uint_as_float(v.y);
        a2 += __uint_as_float(v.z);
        a3 += __uint_as_float(v.w);
    }
    // stand-in for the partial store
    float4 *o = reinterpret_cast<float4 *>(out) + (size_t)blockIdx.x * 256 + threadIdx.x;
    *o = make_float4(a0, a1, a2, a3);
}

int main()
{
    const int n = 4096, row_vec = n / 4;
    const size_t bytes = (size_t)n * n * 4;
    void *buf[4];
    float *out;
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (auto &b : buf) {
        cudaMalloc(&b, bytes);
        cudaMemset(b, 0, bytes);
    }
    cudaMalloc(&out, 64 << 20);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("tile_bytes blocks splits us_per_kernel GB/s\n");
    for (int tile_vec : {32, 64, 128, 256})
        for (int per_sm : {2, 4, 8}) {
            const int tiles = row_vec / tile_vec;
            const int splits = std::max(1, 148 * per_sm / tiles);
            const int blocks = tiles * splits;
            auto body = [&] {
                for (int k = 0; k < 12; ++k) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(blocks);
                    cfg.blockDim = dim3(256);
                    cfg.stream = st;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = 1;
                    cudaLaunchKernelEx(&cfg, rd_cols, (const uint4 *)buf[k & 3], row_vec, n, tile_vec, splits, out);
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
            printf("%6d %6d %4d %7.2f %6.0f\n", tile_vec * 16, blocks, splits, t[4], bytes / (t[4] * 1e-6) / 1e9);
        }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
