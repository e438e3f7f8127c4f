// Internal declarations shared by the translation units of libndb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>

#include "../../include/ndb200.h"

namespace ndb {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
void clear_error();
extern std::atomic<uint64_t> g_launches;

inline int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return NDB_ERR_CUDA;
}

#define NDB_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) return ::ndb::cuda_fail(e_, #call, __FILE__, __LINE__); \
    } while (0)

#define NDB_CHECK_LAUNCH()                                                 \
    do {                                                                   \
        ::ndb::g_launches.fetch_add(1, std::memory_order_relaxed);         \
        cudaError_t e_ = cudaGetLastError();                               \
        if (e_ != cudaSuccess) return ::ndb::cuda_fail(e_, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define NDB_TRY(expr)                  \
    do {                               \
        int st_ = (expr);              \
        if (st_ != NDB_OK) return st_; \
    } while (0)

inline size_t dtype_size(int dt)
{
    switch (dt) {
    case NDB_F32: return 4;
    case NDB_F64: return 8;
    case NDB_I32: return 4;
    case NDB_I64: return 8;
    default: return 0;
    }
}

inline bool dtype_is_float(int dt) { return dt == NDB_F32 || dt == NDB_F64; }

// ---- programmatic dependent launch ---------------------------------------------------------------------------------
// Back-to-back kernels in one stream are spaced by ~3-4 us on these boxes (an EMPTY kernel sustains one launch per
// 3.9 us; benchmarks/micro/launch_gap.cu): more than a third of a 10 us reduction.  Every streaming kernel therefore
// begins with pdl_enter() and is launched through launch_pdl(): the next kernel's blocks are scheduled while this one's
// last wave drains, and block at griddepcontrol.wait until this grid has completed and flushed, so stream order is kept
// exactly (64 MiB cold read, 12 in a row: 13.0 -> 11.75 us per kernel).  A kernel launched without the attribute, or after
// a memcpy / event, sees the usual full serialisation; pdl_enter() is then a no-op.  NDB_NO_PDL=1 turns the attribute off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Kernels that wait for OTHER GPUs (peer-window exchanges) must not release their dependents early: measured on 4 and 8
// GPUs, a step whose exchange kernel let the next Zip's blocks become resident while it was still waiting for its peers
// ran 8-10 % slower (2.99 -> 3.30 ms at N = 4) than the same step with that edge serialised.
__device__ __forceinline__ void pdl_wait_only() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface through cudaGetLastError (NDB_CHECK_LAUNCH)
}
#endif

// number of *input* producers an op consumes (ADD_N: variable, returns -1)
int op_arity(int op);
// number of scalars the op reads
int op_scalars(int op);

int validate_view(const ndb_view *v, const char *what);

}  // namespace ndb

// ---- context ------------------------------------------------------------------
constexpr int NDB_TICKETS = 65536;

struct ndb_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // the one in use
    void *workspace = nullptr;      // device scratch for reduction partials
    size_t workspace_bytes = 0;
    int *dev_flag = nullptr;        // device int: set by kernels that detect a Rust-panic condition
    int *host_flag = nullptr;       // pinned mirror
    void *host_scalar = nullptr;    // pinned 64-byte buffer for scalar results
    int sm_count = 148;
    // peer window (peer.cu): this rank's window and every peer's, mapped through CUDA IPC
    int peer_rank = -1, peer_world = 0;
    bool peer_attached = false;
    size_t peer_slot_bytes = 0;
    void *peer_win[16] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long peer_expected = 0;  // arrivals my counter will have seen once every exchange issued so far is complete
    unsigned long long peer_calls = 0;
    unsigned long long *peer_tile_epoch = nullptr;  // device: exchanges completed per tile counter (reduce.cu's fused axis reductions)
    unsigned *tickets = nullptr;    // NDB_TICKETS zeroed counters: 'last block finishes' hand-over in split reductions (self-resetting)
    // staging buffers for the host-pointer (end-to-end / cblas) entry points
    void *stage[NDB_MAX_PARTS] = {};  // one per Zip producer (ndb_zip_host mirrors up to NDB_MAX_PARTS spans)
    size_t stage_bytes[NDB_MAX_PARTS] = {};
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // pinned bounce buffers for PAGEABLE host memory (the CBLAS boundary hands in ndarray's Vec storage)
    void *bounce[2] = {nullptr, nullptr};
    cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
    void *bounce_out[2] = {nullptr, nullptr};  // second pair: a D2H worker thread drains results while the caller's thread uploads
    cudaEvent_t bounce_out_ev[2] = {nullptr, nullptr};
    // chained scan (scan.cu): ticket + per-chunk packets, and the epoch that marks this launch's packets
    void *scan_chain = nullptr;
    size_t scan_chain_bytes = 0;
    uint32_t scan_epoch = 0;
    // stream-K f64 GEMM (gemm_f64_tma.cu): ticket, one flag per stream-K CTA (all zero between launches) and the partial tiles
    void *streamk = nullptr;
    size_t streamk_bytes = 0;
    // NCCL communicator owned by the library (dist.cu; libnccl is dlopen'ed on first use)
    void *nccl_comm = nullptr;
    int comm_rank = -1, comm_world = 0;
};

namespace ndb {
int ensure_workspace(ndb_ctx *ctx, size_t bytes);
bool stream_is_capturing(ndb_ctx *ctx);  // true while ndb_graph_begin .. ndb_graph_end (or the caller's own capture) is recording ctx->stream
int ensure_stage(ndb_ctx *ctx, int slot, size_t bytes);
// pageable host <-> device through two pinned bounce buffers: a few host threads copy chunk i+1 while the DMA engine
// moves chunk i (cudaMemcpy from pageable memory runs at ~12 GB/s on these boxes, this pipeline at the threads' memcpy rate).
// Ordered after / before work on ctx->stream; returns after the last chunk has been ISSUED (h2d) or has LANDED (d2h).
int prepare_bounce(ndb_ctx *ctx);  // allocate both bounce pairs now (before a second host thread starts using them)
int staged_h2d(ndb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st = nullptr);
int staged_d2h(ndb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, cudaStream_t st = nullptr, int set = 0);
int welford_merge_gathered(ndb_ctx *ctx, int dtype, int world, const int64_t *counts_dev, const void *means_dev, const void *m2s_dev,
                           int64_t n_out, int64_t pitch, double ddof, int want_std, void *out_dev);  // reduce.cu
int write_identity(ndb_ctx *ctx, int dtype, int redop, void *out_dev);  // reduce.cu: what an EMPTY shard contributes to an exchange
struct PeerArgs;
int peer_begin(ndb_ctx *ctx, int blocks, size_t bytes, PeerArgs *out);  // peer.cu
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace ndb
