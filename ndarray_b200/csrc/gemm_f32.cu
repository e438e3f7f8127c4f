// f32 general_mat_mul as error-compensated 3xTF32 on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   a = a_hi + a_lo,  a_hi = tf32(a) (round to nearest),  a_lo = a - a_hi  (exact in f32)
//   A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi      (the dropped A_lo*B_lo term is ~2^-22 relative)
// with every product accumulated in FP32 in tensor memory, which keeps FP32-class accuracy (the reference's
// own bar: max|diff|/max|ref| < 1e-5, crates/numeric-tests/tests/accuracy.rs:146) at tensor-core speed.
//
// Pipeline
//   1. pack: the Zip engine writes A_hi, A_lo as [M][Kp] and B_hi, B_lo as [N][Kp] (K-major, pitch 16-byte
//      aligned) from ANY signed strides (ops NDB_OP_TF32_HI / _LO; a row-major B goes through the
//      tile-transpose kernel).  So C-order, F-order, reversed, stride-2 and broadcast operands all take the
//      tensor-core path, like matrixmultiply::sgemm accepts any strides (impl_linalg.rs:512-528).
//   2. gemm_tf32x3: one 128 x BN output tile per CTA, warp-specialised:
//        warp 0   TMA producer   cp.async.bulk.tensor (128B swizzle) -> 2-stage smem ring, mbarrier tx-count
//        warp 1   MMA issuer     one elected lane: 3 x tcgen05.mma.kind::tf32 per 8-wide k slice, accumulators
//                                in TMEM (BN fp32 columns x 128 lanes); tcgen05.commit frees the smem stage
//        warps 2-5 epilogue      tcgen05.ld 32x32b.x32 -> alpha*acc + beta*C -> global (beta == 0 never reads C)
// No CUTLASS: descriptors and PTX are written out below.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "gemm.h"
#include "zip_plan.h"

extern "C" int ndb_zip(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts);

namespace ndb {

namespace {

constexpr int BM = 128;
constexpr int BK = 32;      // 32 tf32 = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;   // 32 bytes of K per tcgen05.mma
constexpr int THREADS = 192;

template <int BN>
struct Cfg {
    static constexpr int STAGES = BN == 256 ? 2 : 3;  // 96 KB or 64 KB per stage
    static constexpr uint32_t A_BYTES = BM * BK * 4;
    static constexpr uint32_t B_BYTES = BN * BK * 4;
    static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

struct F32Args {
    float *c;
    int64_t m, n, k;
    int64_t rsc, csc;
    float alpha, beta;
    int tiles_m, tiles_n, group_m, chunk_kb;
    int debug;  // tuning probes only: 1 = no TMA traffic after the first ring fill, 2 = no MMAs (results are garbage)
};

// ---- PTX helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

// K-major operand tile as TMA leaves it with CU_TENSOR_MAP_SWIZZLE_128B: rows of 128 bytes, 8-row groups of
// 1024 bytes.  start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 >> 4 | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// kind::tf32, FP32 accumulate, both operands K-major: c_format=F32 (bit 4), a/b_format=TF32 (2 at bits 7, 10),
// N>>3 at bit 17, M>>4 at bit 24
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc()
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Accumulation accuracy.  The tensor core adds into its FP32 accumulator with truncation, so one long K
// chain drifts linearly with K (measured: 9e-9 * K relative, i.e. 6e-5 at K = 8192, back to 1xTF32 quality at
// K = 32768, benchmarks/gemm_accuracy.py).  The kernel therefore accumulates only chunk_kb k-blocks (8: K = 256)
// per TMEM accumulator, double-buffers two accumulators, and lets the epilogue warps drain each finished
// chunk into FP32 registers with round-to-nearest adds while the MMA warp fills the other one.  Error is
// then ~2e-6 independent of K (fp32 class), for one extra TMEM read per 256 k.

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, F32Args p)
{
    using C = Cfg<BN>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-byte aligned tiles
    const uint32_t bars = tiles + STAGES * C::STAGE_BYTES;
    // barrier block: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base address
    const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
    const uint32_t tmem_full_bar = bars + 16 * STAGES, tmem_empty_bar = tmem_full_bar + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + (bars - raw) + 16 * STAGES + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // rasterise in groups of GROUP tile-rows: the ~148 concurrently resident CTAs then cover a roughly square patch
    // of C, so each A/B panel slice fetched from DRAM is reused from L2 by many CTAs of the same wave
    const int GROUP = p.group_m;
    const int per_group = GROUP * p.tiles_n;
    const int grp = blockIdx.x / per_group;
    const int first_m = grp * GROUP;
    const int gsize = min(GROUP, p.tiles_m - first_m);
    const int in_grp = blockIdx.x - grp * per_group;
    const int tile_m = first_m + in_grp % gsize, tile_n = in_grp / gsize;
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    const int num_kb = (int)((p.k + BK - 1) / BK);
    const int CHUNK_KB = p.chunk_kb;
    const int num_chunks = (num_kb + CHUNK_KB - 1) / CHUNK_KB;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full_bar + 8 * b, 1);   // one tcgen05.commit
            mbar_init(tmem_empty_bar + 8 * b, 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // whole warp: allocate two BN-column accumulators, publish the base address through smem
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(empty_bar + 8 * s, ph ^ 1u);
                const uint32_t st = tiles + s * C::STAGE_BYTES;
                if (p.debug == 1 && kb >= STAGES) {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_bar + 8 * s) : "memory");
                    continue;
                }
                mbar_expect_tx(full_bar + 8 * s, C::STAGE_BYTES);
                tma_load_2d(st, &map_a_hi, full_bar + 8 * s, kb * BK, m0);
                tma_load_2d(st + C::A_BYTES, &map_a_lo, full_bar + 8 * s, kb * BK, m0);
                tma_load_2d(st + 2 * C::A_BYTES, &map_b_hi, full_bar + 8 * s, kb * BK, n0);
                tma_load_2d(st + 2 * C::A_BYTES + C::B_BYTES, &map_b_lo, full_bar + 8 * s, kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            constexpr uint32_t idesc = make_idesc<BN>();
            for (int ch = 0; ch < num_chunks; ++ch) {
                const int buf = ch & 1;
                mbar_wait(tmem_empty_bar + 8 * buf, (((uint32_t)ch >> 1) & 1u) ^ 1u);  // drained by the epilogue
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                const int kb_end = min(num_kb, (ch + 1) * CHUNK_KB);
                for (int kb = ch * CHUNK_KB; kb < kb_end; ++kb) {
                    const int s = kb % STAGES;
                    const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                    mbar_wait(full_bar + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t st = tiles + s * C::STAGE_BYTES;
                    const uint64_t da_hi = make_desc(st), da_lo = make_desc(st + C::A_BYTES);
                    const uint64_t db_hi = make_desc(st + 2 * C::A_BYTES), db_lo = make_desc(st + 2 * C::A_BYTES + C::B_BYTES);
                    const uint32_t first = (kb != ch * CHUNK_KB);
                    if (p.debug == 2 && kb > 0) {
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_bar + 8 * s) : "memory");
                        continue;
                    }
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)(k * UMMA_K * 4 >> 4);  // 32 bytes per slice inside the swizzle row
                        tc_mma_tf32(acc, da_lo + adv, db_hi + adv, idesc, first | (uint32_t)(k != 0));
                        tc_mma_tf32(acc, da_hi + adv, db_lo + adv, idesc, 1u);
                        tc_mma_tf32(acc, da_hi + adv, db_hi + adv, idesc, 1u);
                    }
                    tc_commit(empty_bar + 8 * s);  // arrives once these MMAs have finished reading the stage
                }
                tc_commit(tmem_full_bar + 8 * buf);  // this chunk's accumulator is complete
            }
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; thread = one row of the tile
        const int q = warp & 3;
        float sum[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) sum[j] = 0.f;
        for (int ch = 0; ch < num_chunks; ++ch) {
            const int buf = ch & 1;
            mbar_wait(tmem_full_bar + 8 * buf, ((uint32_t)ch >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar + 8 * buf) : "memory");
        }
        const int64_t row = (int64_t)m0 + q * 32 + lane;
        const bool vec_ok = p.csc == 1 && (p.rsc & 3) == 0 && ((uintptr_t)p.c & 15) == 0;
        if (row < p.m) {
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int64_t col0 = (int64_t)n0 + c0;
                if (col0 >= p.n) break;
                float *dst = p.c + row * p.rsc + col0 * p.csc;
                if (vec_ok && col0 + 32 <= p.n) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o;
                        o.x = p.alpha * sum[c0 + j];
                        o.y = p.alpha * sum[c0 + j + 1];
                        o.z = p.alpha * sum[c0 + j + 2];
                        o.w = p.alpha * sum[c0 + j + 3];
                        if (p.beta != 0.f) {
                            const float4 old = *reinterpret_cast<const float4 *>(dst + j);
                            o.x += p.beta * old.x;
                            o.y += p.beta * old.y;
                            o.z += p.beta * old.z;
                            o.w += p.beta * old.w;
                        }
                        *reinterpret_cast<float4 *>(dst + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (col0 + j < p.n) {
                            float o = p.alpha * sum[c0 + j];
                            if (p.beta != 0.f) o += p.beta * dst[j * p.csc];
                            dst[j * p.csc] = o;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)));
    }
}

// ---- host side ------------------------------------------------------------------------------------
// Small problems are launch-bound: four Zip launches to pack A_hi/A_lo/B_hi/B_lo cost more than the GEMM itself
// (256^3: 33 us in all, cuBLAS 13 us).  For operands of at most PACK_SMALL_ELEMS elements one launch packs both operands
// and both halves from any signed strides (blockIdx.y selects the operand; reads follow the operand's unit-stride axis
// when it has one, writes are k-major).  Same split as NDB_OP_TF32_HI/_LO in zip_kernels.cuh.
constexpr int64_t PACK_SMALL_ELEMS = 1 << 18;  // 512 x 512: above that the Zip engine's vector kernels win (1024^3: 49.6 vs 53.6 us)

struct PackSmallArgs {
    const float *src[2];
    int64_t rows[2], rs[2], cs[2];  // element (r, k) of operand o at src[o][r * rs[o] + k * cs[o]]
    float *hi[2], *lo[2];
    int64_t k, kp;
};

__global__ void __launch_bounds__(256) pack_tf32_small(PackSmallArgs a)
{
    const int o = blockIdx.y;
    const int64_t rows = a.rows[o], total = rows * a.kp;
    // walk the source along its denser axis so that neighbouring threads read neighbouring addresses
    const bool k_fast = llabs(a.cs[o]) <= llabs(a.rs[o]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r, k;
        if (k_fast) {
            r = i / a.kp;
            k = i - r * a.kp;
        } else {
            k = i / rows;
            r = i - k * rows;
        }
        float h = 0.f, l = 0.f;
        if (k < a.k) {
            const float x = a.src[o][r * a.rs[o] + k * a.cs[o]];
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
            h = __uint_as_float(hb);
            l = (fabsf(x) <= 3.4028234664e38f) ? x - h : 0.f;
        }
        a.hi[o][r * a.kp + k] = h;
        a.lo[o][r * a.kp + k] = l;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [rows][k] f32, row pitch `pitch` elements (multiple of 4), box = BK x box_rows, 128-byte swizzle, OOB -> 0
int make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t k, int64_t pitch, int box_rows)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("gemm_f32: cuTensorMapEncodeTiled is not available from this driver");
        return NDB_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_f32: cuTensorMapEncodeTiled failed with %d", (int)r);
        return NDB_ERR_CUDA;
    }
    return NDB_OK;
}

ndb_view view2(const void *ptr, int64_t d0, int64_t d1, int64_t s0, int64_t s1)
{
    ndb_view v;
    memset(&v, 0, sizeof(v));
    v.ptr = const_cast<void *>(ptr);
    v.ndim = 2;
    v.dtype = NDB_F32;
    v.dim[0] = d0;
    v.dim[1] = d1;
    v.stride[0] = s0;
    v.stride[1] = s1;
    return v;
}

template <int BN>
int launch(ndb_ctx *ctx, const GemmArgs &g, float alpha, float beta, const float *a_hi, const float *a_lo, const float *b_hi,
           const float *b_lo, int64_t kp)
{
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    NDB_TRY(make_map(&ma_hi, a_hi, g.m, g.k, kp, BM));
    NDB_TRY(make_map(&ma_lo, a_lo, g.m, g.k, kp, BM));
    NDB_TRY(make_map(&mb_hi, b_hi, g.n, g.k, kp, BN));
    NDB_TRY(make_map(&mb_lo, b_lo, g.n, g.k, kp, BN));
    F32Args p;
    p.c = (float *)g.c;
    p.m = g.m; p.n = g.n; p.k = g.k;
    p.rsc = g.rsc; p.csc = g.csc;
    p.alpha = alpha; p.beta = beta;
    p.tiles_m = (int)((g.m + BM - 1) / BM);
    p.tiles_n = (int)((g.n + BN - 1) / BN);
    p.group_m = getenv("NDB_GEMM_GROUP_M") ? std::max(1, atoi(getenv("NDB_GEMM_GROUP_M"))) : 12;
    p.debug = getenv("NDB_GEMM_DEBUG") ? atoi(getenv("NDB_GEMM_DEBUG")) : 0;
    p.chunk_kb = getenv("NDB_GEMM_CHUNK_KB") ? std::max(1, atoi(getenv("NDB_GEMM_CHUNK_KB"))) : 8;
    // the opt-in is per DEVICE (a process may hold one context per GPU); setting it twice from racing threads is harmless
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
        NDB_CUDA(cudaFuncSetAttribute(gemm_tf32x3<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BN>::SMEM_BYTES));
        attr_done.fetch_or(dev_bit, std::memory_order_release);
    }
    const unsigned blocks = (unsigned)(p.tiles_m * p.tiles_n);
    gemm_tf32x3<BN><<<blocks, THREADS, Cfg<BN>::SMEM_BYTES, ctx->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

}  // namespace

bool gemm_f32_eligible(const GemmArgs &g)
{
    if (getenv("NDB_GEMM_F32_GENERIC")) return false;
    if (g.m < 64 || g.n < 64 || g.k < 32) return false;
    {
        // below this many multiply-adds the FFMA tile kernel (gemm_any) is used instead; NDB_GEMM_F32_MIN_CUBE = edge of the cube
        const double edge = getenv("NDB_GEMM_F32_MIN_CUBE") ? atof(getenv("NDB_GEMM_F32_MIN_CUBE")) : 128.0;
        if ((double)g.m * (double)g.n * (double)g.k < edge * edge * edge) return false;
    }
    if (g.m > (1 << 30) || g.n > (1 << 30) || g.k > (1 << 30)) return false;
    return true;
}

int gemm_f32_launch(ndb_ctx *ctx, const GemmArgs &g0, float alpha, float beta)
{
    GemmArgs g = g0;
    // write C along its denser axis: a column-major C is computed as C^T = B^T A^T
    if (llabs(g.rsc) < llabs(g.csc)) {
        GemmArgs t = g;
        t.a = g.b; t.rsa = g.csb; t.csa = g.rsb;
        t.b = g.a; t.rsb = g.csa; t.csb = g.rsa;
        t.rsc = g.csc; t.csc = g.rsc;
        t.m = g.n; t.n = g.m;
        g = t;
    }
    const int64_t kp = (g.k + 3) & ~(int64_t)3;
    const size_t a_elems = (size_t)g.m * kp, b_elems = (size_t)g.n * kp;
    float *ws = nullptr;
    NDB_CUDA(cudaMallocAsync((void **)&ws, (2 * a_elems + 2 * b_elems) * sizeof(float), ctx->stream));
    float *a_hi = ws, *a_lo = ws + a_elems, *b_hi = ws + 2 * a_elems, *b_lo = ws + 2 * a_elems + b_elems;
    int st = NDB_OK;
    if ((int64_t)a_elems <= PACK_SMALL_ELEMS && (int64_t)b_elems <= PACK_SMALL_ELEMS && !getenv("NDB_GEMM_F32_ZIP_PACK")) {
        PackSmallArgs pa;
        pa.src[0] = (const float *)g.a; pa.rows[0] = g.m; pa.rs[0] = g.rsa; pa.cs[0] = g.csa; pa.hi[0] = a_hi; pa.lo[0] = a_lo;
        pa.src[1] = (const float *)g.b; pa.rows[1] = g.n; pa.rs[1] = g.csb; pa.cs[1] = g.rsb; pa.hi[1] = b_hi; pa.lo[1] = b_lo;
        pa.k = g.k; pa.kp = kp;
        const int64_t most = (int64_t)std::max(a_elems, b_elems);
        const unsigned bx = (unsigned)std::min<int64_t>((most + 255) / 256, (int64_t)ctx->sm_count * 8);
        pack_tf32_small<<<dim3(bx, 2, 1), 256, 0, ctx->stream>>>(pa);
        NDB_CHECK_LAUNCH();
    } else {
        // pack + split through the Zip engine: out[M][Kp] <- op(A),  out[N][Kp] <- op(B^T)
        const ndb_view va = view2(g.a, g.m, g.k, g.rsa, g.csa);
        const ndb_view vbt = view2(g.b, g.n, g.k, g.csb, g.rsb);
        const struct { float *dst; const ndb_view *src; int64_t rows; int op; } jobs[4] = {
            {a_hi, &va, g.m, NDB_OP_TF32_HI}, {a_lo, &va, g.m, NDB_OP_TF32_LO},
            {b_hi, &vbt, g.n, NDB_OP_TF32_HI}, {b_lo, &vbt, g.n, NDB_OP_TF32_LO}};
        for (int j = 0; j < 4 && st == NDB_OK; ++j) {
            ndb_view parts[2] = {view2(jobs[j].dst, jobs[j].rows, g.k, kp, 1), *jobs[j].src};
            st = ndb_zip(ctx, jobs[j].op, nullptr, 2, parts);
        }
    }
    if (st == NDB_OK) {
        // 128-wide tiles with a 3-stage ring measured faster than 256-wide with 2 stages (248 vs 221 TFLOP/s at 8192^3)
        st = launch<128>(ctx, g, alpha, beta, a_hi, a_lo, b_hi, b_lo, kp);
    }
    cudaFreeAsync(ws, ctx->stream);
    return st;
}

}  // namespace ndb
