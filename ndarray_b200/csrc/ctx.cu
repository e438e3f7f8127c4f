// Context, memory, events and error plumbing of libndb200.
#include <stdarg.h>

#include <algorithm>
#include <cstdlib>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.h"

namespace ndb {

static thread_local char tl_error[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
}

void clear_error() { tl_error[0] = 0; }

bool pdl_enabled()
{
    static const bool on = getenv("NDB_NO_PDL") == nullptr;
    return on;
}

int op_arity(int op)
{
    switch (op) {
    case NDB_OP_FILL: return 0;
    case NDB_OP_COPY:
    case NDB_OP_ADD_S:
    case NDB_OP_SUB_S:
    case NDB_OP_MUL_S:
    case NDB_OP_DIV_S:
    case NDB_OP_REM_S:
    case NDB_OP_RSUB_S:
    case NDB_OP_RDIV_S:
    case NDB_OP_NEG:
    case NDB_OP_ABS:
    case NDB_OP_TF32_HI:
    case NDB_OP_TF32_LO:
    case NDB_OP_MATH:
    case NDB_OP_CLAMP:
    case NDB_OP_SQRT: return 1;
    case NDB_OP_ADD:
    case NDB_OP_SUB:
    case NDB_OP_MUL:
    case NDB_OP_DIV:
    case NDB_OP_REM:
    case NDB_OP_AXPY:
    case NDB_OP_MIN:
    case NDB_OP_MAX:
    case NDB_OP_EQ:
    case NDB_OP_NE:
    case NDB_OP_LT:
    case NDB_OP_LE:
    case NDB_OP_GT:
    case NDB_OP_GE:
    case NDB_OP_SCALE_ADD: return 2;
    case NDB_OP_MUL_ADD: return 3;
    case NDB_OP_ADD_N: return -1;
    default: return -2;
    }
}

int op_scalars(int op)
{
    switch (op) {
    case NDB_OP_FILL:
    case NDB_OP_ADD_S:
    case NDB_OP_SUB_S:
    case NDB_OP_MUL_S:
    case NDB_OP_DIV_S:
    case NDB_OP_REM_S:
    case NDB_OP_RSUB_S:
    case NDB_OP_RDIV_S:
    case NDB_OP_AXPY: return 1;
    case NDB_OP_MATH:
    case NDB_OP_CLAMP:
    case NDB_OP_SCALE_ADD: return 2;
    default: return 0;
    }
}

int validate_view(const ndb_view *v, const char *what)
{
    if (!v) {
        set_error("%s: null view", what);
        return NDB_ERR_INVALID_ARG;
    }
    if (v->ndim < 0 || v->ndim > NDB_MAX_NDIM) {
        set_error("%s: ndim %d outside 0..%d", what, v->ndim, NDB_MAX_NDIM);
        return NDB_ERR_UNSUPPORTED;
    }
    if (dtype_size(v->dtype) == 0) {
        set_error("%s: unknown dtype %d", what, v->dtype);
        return NDB_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < v->ndim; ++i) {
        if (v->dim[i] < 0) {
            set_error("%s: negative axis length", what);
            return NDB_ERR_INVALID_ARG;
        }
    }
    return NDB_OK;
}

bool stream_is_capturing(ndb_ctx *ctx)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone;
}

int ensure_workspace(ndb_ctx *ctx, size_t bytes)
{
    if (ctx->workspace_bytes >= bytes) return NDB_OK;
    // Growing must not race with kernels still using the old buffer.
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->workspace) NDB_CUDA(cudaFree(ctx->workspace));
    ctx->workspace = nullptr;
    ctx->workspace_bytes = 0;
    size_t want = bytes < (size_t(8) << 20) ? (size_t(8) << 20) : bytes;
    NDB_CUDA(cudaMalloc(&ctx->workspace, want));
    ctx->workspace_bytes = want;
    return NDB_OK;
}

namespace {

constexpr size_t BOUNCE_BYTES = size_t(16) << 20;

int ensure_bounce(ndb_ctx *ctx)
{
    for (int i = 0; i < 2; ++i) {
        if (!ctx->bounce[i]) NDB_CUDA(cudaHostAlloc(&ctx->bounce[i], BOUNCE_BYTES, cudaHostAllocDefault));
        if (!ctx->bounce_ev[i]) NDB_CUDA(cudaEventCreateWithFlags(&ctx->bounce_ev[i], cudaEventDisableTiming));
        if (!ctx->bounce_out[i]) NDB_CUDA(cudaHostAlloc(&ctx->bounce_out[i], BOUNCE_BYTES, cudaHostAllocDefault));
        if (!ctx->bounce_out_ev[i]) NDB_CUDA(cudaEventCreateWithFlags(&ctx->bounce_out_ev[i], cudaEventDisableTiming));
    }
    return NDB_OK;
}

// memcpy split over a few threads (one thread moves ~10 GB/s; the DMA engine wants ~50).  The helpers are a persistent
// pool per calling thread: spawning threads per 16 MiB chunk cost a third of the copy itself.
class CopyPool {
public:
    explicit CopyPool(int helpers)
    {
        for (int i = 0; i < helpers; ++i) th_.emplace_back([this, i] { work(i + 1); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int parts() const { return (int)th_.size() + 1; }
    void run(void *dst, const void *src, size_t bytes)
    {
        const int n = parts();
        const size_t per = ((bytes + n - 1) / n + 4095) & ~size_t(4095);
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = (char *)dst;
            src_ = (const char *)src;
            bytes_ = bytes;
            per_ = per;
            pending_ = n - 1;
            ++gen_;
        }
        cv_.notify_all();
        memcpy(dst, src, std::min(per, bytes));
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    void work(int id)
    {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            char *d = dst_;
            const char *s = src_;
            const size_t bytes = bytes_, per = per_;
            lk.unlock();
            const size_t lo = per * (size_t)id;
            if (lo < bytes) memcpy(d + lo, s + lo, std::min(per, bytes - lo));
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, per_ = 0;
};

void parallel_memcpy(void *dst, const void *src, size_t bytes)
{
    // two of these run side by side in the pipelined CBLAS path (upload on the caller's thread, download on a worker), so a
    // team is 3/8 of the hardware threads: 6 of 16 measured faster than 8 of 16 (sgemm 8192^3 end to end 28.2 -> 23.7 ms)
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<unsigned>(8, std::max<unsigned>(1, hw * 3 / 8));
    if (getenv("NDB_COPY_THREADS")) nt = std::max(1, atoi(getenv("NDB_COPY_THREADS")));
    if (bytes < (size_t(1) << 20) || nt == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    thread_local std::unique_ptr<CopyPool> pool;
    if (!pool || pool->parts() != nt) pool.reset(new CopyPool(nt - 1));
    pool->run(dst, src, bytes);
}

}  // namespace

int prepare_bounce(ndb_ctx *ctx) { return ensure_bounce(ctx); }

int staged_h2d(ndb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st)
{
    if (!st) st = ctx->stream;
    if (bytes < (size_t(4) << 20) || getenv("NDB_NO_BOUNCE")) {
        NDB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
        return NDB_OK;
    }
    NDB_TRY(ensure_bounce(ctx));
    size_t done = 0;
    for (int i = 0; done < bytes; ++i) {
        const int b = i & 1;
        const size_t len = std::min(BOUNCE_BYTES, bytes - done);
        if (i >= 2) NDB_CUDA(cudaEventSynchronize(ctx->bounce_ev[b]));  // the DMA that last read this buffer is done
        parallel_memcpy(ctx->bounce[b], (const char *)src_host + done, len);
        NDB_CUDA(cudaMemcpyAsync((char *)dst_dev + done, ctx->bounce[b], len, cudaMemcpyHostToDevice, st));
        NDB_CUDA(cudaEventRecord(ctx->bounce_ev[b], st));
        done += len;
    }
    // the bounce buffers are reused by the next call: wait until the DMA engine has drained them
    NDB_CUDA(cudaEventSynchronize(ctx->bounce_ev[0]));
    NDB_CUDA(cudaEventSynchronize(ctx->bounce_ev[1]));
    return NDB_OK;
}

int staged_d2h(ndb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, cudaStream_t st, int set)
{
    if (!st) st = ctx->stream;
    if (bytes < (size_t(4) << 20) || getenv("NDB_NO_BOUNCE")) {
        NDB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        NDB_CUDA(cudaStreamSynchronize(st));
        return NDB_OK;
    }
    NDB_TRY(ensure_bounce(ctx));
    void **buf = set ? ctx->bounce_out : ctx->bounce;
    cudaEvent_t *ev = set ? ctx->bounce_out_ev : ctx->bounce_ev;
    const size_t chunks = (bytes + BOUNCE_BYTES - 1) / BOUNCE_BYTES;
    auto issue = [&](size_t i) -> cudaError_t {
        const size_t off = i * BOUNCE_BYTES, len = std::min(BOUNCE_BYTES, bytes - off);
        cudaError_t e = cudaMemcpyAsync(buf[i & 1], (const char *)src_dev + off, len, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(ev[i & 1], st);
        return e;
    };
    NDB_CUDA(issue(0));
    for (size_t i = 0; i < chunks; ++i) {
        if (i + 1 < chunks) NDB_CUDA(issue(i + 1));  // next chunk's DMA overlaps this chunk's host copy
        NDB_CUDA(cudaEventSynchronize(ev[i & 1]));
        const size_t off = i * BOUNCE_BYTES, len = std::min(BOUNCE_BYTES, bytes - off);
        parallel_memcpy((char *)dst_host + off, buf[i & 1], len);
    }
    return NDB_OK;
}

int ensure_stage(ndb_ctx *ctx, int slot, size_t bytes)
{
    static_assert(sizeof(ctx->stage) / sizeof(ctx->stage[0]) == NDB_MAX_PARTS, "one staging slot per Zip producer");
    if (slot < 0 || slot >= NDB_MAX_PARTS) {
        set_error("ensure_stage: slot %d outside 0..%d", slot, NDB_MAX_PARTS - 1);
        return NDB_ERR_INVALID_ARG;
    }
    if (ctx->stage_bytes[slot] >= bytes) return NDB_OK;
    NDB_CUDA(cudaDeviceSynchronize());
    if (ctx->stage[slot]) NDB_CUDA(cudaFree(ctx->stage[slot]));
    ctx->stage[slot] = nullptr;
    ctx->stage_bytes[slot] = 0;
    NDB_CUDA(cudaMalloc(&ctx->stage[slot], bytes));
    ctx->stage_bytes[slot] = bytes;
    return NDB_OK;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

const char *ndb_version(void) { return "ndb200 0.1.0 (sm_100a)"; }

const char *ndb_last_error(void) { return tl_error; }

uint64_t ndb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int ndb_device_count(int *count)
{
    if (!count) return NDB_ERR_INVALID_ARG;
    *count = 0;
    NDB_CUDA(cudaGetDeviceCount(count));
    return NDB_OK;
}

int ndb_ctx_create(int device, ndb_ctx **out)
{
    if (!out) return NDB_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    NDB_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        set_error("ndb_ctx_create: device %d not in 0..%d (no CPU fallback exists)", device, n);
        return NDB_ERR_CUDA;
    }
    NDB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NDB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("ndb_ctx_create: device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                  prop.minor);
        return NDB_ERR_CUDA;
    }
    ndb_ctx *ctx = new ndb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    NDB_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    NDB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    NDB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    NDB_CUDA(cudaMalloc((void **)&ctx->dev_flag, sizeof(int)));
    NDB_CUDA(cudaMemset(ctx->dev_flag, 0, sizeof(int)));
    NDB_CUDA(cudaMallocHost((void **)&ctx->host_flag, sizeof(int)));
    NDB_CUDA(cudaMallocHost(&ctx->host_scalar, 64));
    *ctx->host_flag = 0;
    {
        // keep freed blocks cached in the pool instead of returning them to the OS at every sync
        cudaMemPool_t pool;
        NDB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;
        NDB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    if (cudaMalloc((void **)&ctx->tickets, NDB_TICKETS * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(ctx->tickets, 0, NDB_TICKETS * sizeof(unsigned)) != cudaSuccess) {
        set_error("ctx_create: cannot allocate the ticket counters");
        return NDB_ERR_CUDA;
    }
    int st = ensure_workspace(ctx, size_t(8) << 20);
    if (st != NDB_OK) return st;
    *out = ctx;
    return NDB_OK;
}

int ndb_ctx_destroy(ndb_ctx *ctx)
{
    if (!ctx) return NDB_OK;
    ndb_peer_window_destroy(ctx);
    ndb_comm_destroy(ctx);
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->workspace) cudaFree(ctx->workspace);
    for (int i = 0; i < NDB_MAX_PARTS; ++i)
        if (ctx->stage[i]) cudaFree(ctx->stage[i]);
    if (ctx->tickets) cudaFree(ctx->tickets);
    if (ctx->scan_chain) cudaFree(ctx->scan_chain);
    if (ctx->streamk) cudaFree(ctx->streamk);
    if (ctx->dev_flag) cudaFree(ctx->dev_flag);
    if (ctx->host_flag) cudaFreeHost(ctx->host_flag);
    if (ctx->host_scalar) cudaFreeHost(ctx->host_scalar);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    for (int i = 0; i < 2; ++i) {
        if (ctx->bounce[i]) cudaFreeHost(ctx->bounce[i]);
        if (ctx->bounce_ev[i]) cudaEventDestroy(ctx->bounce_ev[i]);
        if (ctx->bounce_out[i]) cudaFreeHost(ctx->bounce_out[i]);
        if (ctx->bounce_out_ev[i]) cudaEventDestroy(ctx->bounce_out_ev[i]);
    }
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    delete ctx;
    return NDB_OK;
}

int ndb_ctx_set_stream(ndb_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return NDB_ERR_INVALID_ARG;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return NDB_OK;
}

void *ndb_ctx_stream(ndb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int ndb_sync(ndb_ctx *ctx)
{
    if (!ctx) return NDB_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return NDB_OK;
}

int ndb_alloc(ndb_ctx *ctx, size_t bytes, void **dptr)
{
    if (!ctx || !dptr) return NDB_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    *dptr = nullptr;
    if (bytes == 0) bytes = 16;  // a valid, dereferenceable-by-nobody pointer for empty arrays
    // Stream-ordered allocation from the device's default pool (kept warm, see ndb_ctx_create): a fresh
    // output per map_collect / sum_axis costs microseconds, not a cudaMalloc + implicit sync.
    NDB_CUDA(cudaMallocAsync(dptr, bytes, ctx->stream));
    return NDB_OK;
}

int ndb_free(ndb_ctx *ctx, void *dptr)
{
    if (!ctx) return NDB_ERR_INVALID_ARG;
    if (!dptr) return NDB_OK;
    DeviceGuard g(ctx->device);
    NDB_CUDA(cudaFreeAsync(dptr, ctx->stream));
    return NDB_OK;
}

int ndb_memset(ndb_ctx *ctx, void *dptr, int byte, size_t bytes)
{
    if (!ctx) return NDB_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    if (bytes) NDB_CUDA(cudaMemsetAsync(dptr, byte, bytes, ctx->stream));
    return NDB_OK;
}

int ndb_h2d(ndb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes)
{
    if (!ctx) return NDB_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    if (bytes) NDB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return NDB_OK;
}

int ndb_d2h(ndb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes)
{
    if (!ctx) return NDB_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    if (bytes) NDB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return NDB_OK;
}

int ndb_d2d(ndb_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes)
{
    if (!ctx) return NDB_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    if (bytes) NDB_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return NDB_OK;
}

int ndb_host_alloc(size_t bytes, void **hptr)
{
    if (!hptr) return NDB_ERR_INVALID_ARG;
    *hptr = nullptr;
    NDB_CUDA(cudaMallocHost(hptr, bytes ? bytes : 16));
    return NDB_OK;
}

int ndb_host_free(void *hptr)
{
    if (hptr) NDB_CUDA(cudaFreeHost(hptr));
    return NDB_OK;
}

// ---- CUDA graphs: replay a fixed sequence of calls without per-launch host cost ---------------------------------------
int ndb_graph_begin(ndb_ctx *ctx)
{
    if (!ctx) return NDB_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    NDB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    return NDB_OK;
}

int ndb_graph_end(ndb_ctx *ctx, void **graph_out)
{
    if (!ctx || !graph_out) return NDB_ERR_INVALID_ARG;
    *graph_out = nullptr;
    DeviceGuard g(ctx->device);
    cudaGraph_t graph = nullptr;
    NDB_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
    *graph_out = (void *)exec;
    return NDB_OK;
}

int ndb_graph_launch(ndb_ctx *ctx, void *graph)
{
    if (!ctx || !graph) return NDB_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    NDB_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph, ctx->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return NDB_OK;
}

int ndb_graph_destroy(void *graph)
{
    if (graph) NDB_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph));
    return NDB_OK;
}

int ndb_event_create(void **ev)
{
    if (!ev) return NDB_ERR_INVALID_ARG;
    cudaEvent_t e;
    NDB_CUDA(cudaEventCreate(&e));
    *ev = (void *)e;
    return NDB_OK;
}

int ndb_event_destroy(void *ev)
{
    if (ev) NDB_CUDA(cudaEventDestroy((cudaEvent_t)ev));
    return NDB_OK;
}

int ndb_event_record(ndb_ctx *ctx, void *ev)
{
    if (!ctx || !ev) return NDB_ERR_INVALID_ARG;
    DeviceGuard g(ctx->device);
    NDB_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx->stream));
    return NDB_OK;
}

int ndb_event_elapsed_ms(void *start, void *stop, float *ms)
{
    if (!start || !stop || !ms) return NDB_ERR_INVALID_ARG;
    NDB_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
    NDB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return NDB_OK;
}

}  // extern "C"
