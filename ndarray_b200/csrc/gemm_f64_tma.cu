// f64 general_mat_mul, TMA-fed variant of gemm_f64.cu's DMMA kernel (same tiling, same arithmetic, same results).
//
// Why it exists: with per-thread cp.async staging the FP64 tensor pipe tops out at 33.1 TFLOP/s; the same kernel with
// the global->shared copies removed runs at 36.0 (benchmarks/build_variant.sh, NDB_F64_DEBUG=1), i.e. the LDGSTS traffic
// through the SM's load/store path costs 8 % no matter how it is scheduled between the DMMAs.  Here one producer thread
// hands whole stages to the TMA engine (cp.async.bulk.tensor, completion on an mbarrier), so the eight DMMA warps issue
// nothing but LDS.64 + DMMA, and stages are handed over through full/empty mbarriers instead of a CTA-wide barrier.
//
// Shared-memory layout without padding: TMA cannot pad rows, and a plain [row][BK] tile would put the four rows a
// half-warp reads into the same banks.  Both operand orientations are therefore described to TMA as 3-D tensors whose
// innermost extent is 4 doubles (32 B):
//   k-contiguous operand (element (r,k) at base[r*ld + k]):  dims (4 k, rows, K/4), strides (8 B, ld*8 B, 32 B),
//       box (4, 128, BK/4)  -> smem [k/4][row][k%4]:   a half-warp's 4 rows x 4 k are 128 contiguous bytes;
//   row-contiguous operand (element (r,k) at base[k*ld + r]): dims (4 r, K, rows/4), strides (8 B, ld*8 B, 32 B),
//       box (4, BK, 128/4)  -> smem [row/4][k][row%4]: a half-warp's 4 k x 4 rows are 128 contiguous bytes.
// Every fragment LDS.64 is conflict-free, tile overhang is zero-filled by TMA's bounds check.  The split of an axis into
// (4, n/4) needs that axis to be a multiple of 4 (K for k-contiguous operands, the row count for row-contiguous ones);
// other shapes take the cp.async kernel.
//
// Reference semantics: src/linalg/impl_linalg.rs:633-643 (general_mat_mul), :512-546 (dgemm call with arbitrary strides).
#include <cuda.h>

#include <cstdlib>

#include "gemm.h"

namespace ndb {

namespace {

constexpr int BK = 32, STAGES = 3;

// Two tile shapes: 128 x 128 (8 DMMA warps as 2 x 4, 64 x 32 warp tiles, one CTA per SM) for problems that fill the
// machine, and 64 x 64 (4 DMMA warps as 2 x 2, 32 x 32 warp tiles, two CTAs per SM) when 128 x 128 tiles would leave
// SMs idle (a 1024^2 output is only 64 big tiles on 148 SMs).
template <int TILE>
struct Shape {
    static constexpr int BM = TILE, BN = TILE;
    static constexpr int WARPS_M = 2, WARPS_N = TILE == 128 ? 4 : 2;
    static constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N, THREADS = 32 * (CONSUMER_WARPS + 1);
    static constexpr int MI = BM / (WARPS_M * 8), NJ = BN / (WARPS_N * 8);  // DMMA tiles per warp
    static constexpr int A_STAGE = BM * BK, B_STAGE = BN * BK;              // doubles
    static constexpr uint32_t STAGE_BYTES = (A_STAGE + B_STAGE) * sizeof(double);
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 2 * STAGES * sizeof(uint64_t);
    static constexpr int MIN_CTAS = TILE == 128 ? 1 : 2;
};

struct TmaArgs {
    double *c;
    int64_t m, n, k, ldc;
    double alpha, beta;
    int tiles_m, tiles_n;
    int csize;  // thread-block cluster size along K (1 = every CTA owns a whole output tile)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_cluster_f64(uint32_t local_addr, uint32_t rank)
{
    uint32_t remote;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
    return v;
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// element (row, k) of a staged ROWS x BK tile, k = kk + t with kk a multiple of 4
template <bool K_CONTIG, int ROWS>
__device__ __forceinline__ double frag(const double *st, int row, int kk, int t)
{
    if constexpr (K_CONTIG) return st[((kk >> 2) * ROWS + row) * 4 + t];
    else return st[((row >> 2) * BK + kk + t) * 4 + (row & 3)];
}

// linear tile id -> (tile_m, tile_n): groups of 8 tile-rows, so that concurrently resident blocks share A and B panels in L2
__device__ __forceinline__ void raster(const TmaArgs &p, int bid, int &tile_m, int &tile_n)
{
    constexpr int GROUP = 8;
    const int per_group = GROUP * p.tiles_n;
    const int grp = bid / per_group;
    const int first_m = grp * GROUP;
    const int gsize = min(GROUP, p.tiles_m - first_m);
    const int in_grp = bid - grp * per_group;
    tile_m = first_m + in_grp % gsize;
    tile_n = in_grp / gsize;
}

// one staged BK slab into the warp's accumulators; the stage goes back to the producer as soon as its last fragments are
// in registers
template <bool A_KC, bool B_KC, int TILE>
__device__ __forceinline__ void consume_stage(const double *as, const double *bs, double (&acc)[Shape<TILE>::MI][Shape<TILE>::NJ][2], int wm,
                                              int wn, int g, int t, int lane, uint32_t empty_bar)
{
    using S = Shape<TILE>;
    constexpr int BM = S::BM, BN = S::BN, MI = S::MI, NJ = S::NJ, WTM = MI * 8, WTN = NJ * 8;
    double af[2][MI], bf[2][NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i) af[0][i] = frag<A_KC, BM>(as, wm * WTM + i * 8 + g, 0, t);
#pragma unroll
    for (int j = 0; j < NJ; ++j) bf[0][j] = frag<B_KC, BN>(bs, wn * WTN + j * 8 + g, 0, t);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
        const int cur = (kk >> 2) & 1, nxt = cur ^ 1;
        if (kk + 4 < BK) {
#pragma unroll
            for (int i = 0; i < MI; ++i) af[nxt][i] = frag<A_KC, BM>(as, wm * WTM + i * 8 + g, kk + 4, t);
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[nxt][j] = frag<B_KC, BN>(bs, wn * WTN + j * 8 + g, kk + 4, t);
        } else {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar);
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
    }
}

// epilogue: thread owns rows i*8+g, column pairs j*8+2t; alpha/beta fused, beta == 0 never reads C
template <int TILE>
__device__ __forceinline__ void epilogue(const TmaArgs &p, const double (&acc)[Shape<TILE>::MI][Shape<TILE>::NJ][2], int64_t m0, int64_t n0,
                                         int wm, int wn, int g, int t)
{
    using S = Shape<TILE>;
    constexpr int MI = S::MI, NJ = S::NJ, WTM = MI * 8, WTN = NJ * 8;
    const bool even_ld = (p.ldc & 1) == 0;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int64_t row = m0 + wm * WTM + i * 8 + g;
        if (row >= p.m) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int64_t col = n0 + wn * WTN + j * 8 + 2 * t;
            if (col >= p.n) continue;
            double *dst = p.c + row * p.ldc + col;
            double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
            if (col + 1 < p.n && even_ld) {
                if (p.beta != 0.0) {
                    const double2 old = *reinterpret_cast<const double2 *>(dst);
                    v0 = fma(p.beta, old.x, v0);
                    v1 = fma(p.beta, old.y, v1);
                }
                *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
            } else {
                if (p.beta != 0.0) v0 = fma(p.beta, dst[0], v0);
                dst[0] = v0;
                if (col + 1 < p.n) {
                    if (p.beta != 0.0) v1 = fma(p.beta, dst[1], v1);
                    dst[1] = v1;
                }
            }
        }
    }
}

template <bool A_KC, bool B_KC, int TILE>
__global__ void __launch_bounds__(Shape<TILE>::THREADS, Shape<TILE>::MIN_CTAS)
gemm_f64_tma(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TmaArgs p)
{
    using S = Shape<TILE>;
    constexpr int BM = S::BM, BN = S::BN, MI = S::MI, NJ = S::NJ, CONSUMER_WARPS = S::CONSUMER_WARPS;
    constexpr int A_STAGE = S::A_STAGE, B_STAGE = S::B_STAGE;
    constexpr uint32_t STAGE_BYTES = S::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
    double *tiles = reinterpret_cast<double *>(smem_raw + pad);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + pad + (size_t)STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    constexpr bool SPLIT = TILE == 64;  // split-K clusters exist for the small tile only: the big one keeps its registers
    const int csize = SPLIT ? p.csize : 1;
    uint32_t crank = 0;
    if (SPLIT && csize > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    int tile_m, tile_n;
    raster(p, SPLIT ? (int)blockIdx.x / csize : (int)blockIdx.x, tile_m, tile_n);
    const int64_t m0 = (int64_t)tile_m * BM, n0 = (int64_t)tile_n * BN;
    // split-K over the cluster: rank r accumulates k-tiles [kt0, kt0 + ktiles)
    const int all_ktiles = (int)((p.k + BK - 1) / BK);
    const int kt0 = SPLIT ? (int)((int64_t)all_ktiles * crank / csize) : 0;
    const int ktiles = SPLIT ? (int)((int64_t)all_ktiles * (crank + 1) / csize) - kt0 : all_ktiles;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);                 // the producer's expect_tx arrival
            mbar_init(empty0 + 8 * s, CONSUMER_WARPS);   // one arrival per DMMA warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const bool producer = warp == CONSUMER_WARPS;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / S::WARPS_N, wn = warp % S::WARPS_N;
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (producer) {
        // ---- producer: one thread feeds the TMA engine ----
        if (lane == 0) {
            for (int kt = 0; kt < ktiles; ++kt) {
                const int s = kt % STAGES;
                const int gk = kt0 + kt;
                if (kt >= STAGES) mbar_wait(empty0 + 8 * s, (uint32_t)(kt / STAGES - 1) & 1u);
                const uint32_t a_dst = smem_u32(tiles + (size_t)s * (A_STAGE + B_STAGE));
                const uint32_t b_dst = a_dst + A_STAGE * sizeof(double);
                mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
                if constexpr (A_KC) tma_load_3d(a_dst, &map_a, full0 + 8 * s, 0, (int)m0, gk * (BK / 4));
                else tma_load_3d(a_dst, &map_a, full0 + 8 * s, 0, gk * BK, (int)(m0 >> 2));
                if constexpr (B_KC) tma_load_3d(b_dst, &map_b, full0 + 8 * s, 0, (int)n0, gk * (BK / 4));
                else tma_load_3d(b_dst, &map_b, full0 + 8 * s, 0, gk * BK, (int)(n0 >> 2));
            }
        }
    } else
    // ---- consumers: WARPS_M x WARPS_N warps, (MI*8) x (NJ*8) warp tile ----
    for (int kt = 0; kt < ktiles; ++kt) {
        const int s = kt % STAGES;
        mbar_wait(full0 + 8 * s, (uint32_t)(kt / STAGES) & 1u);
        const double *as = tiles + (size_t)s * (A_STAGE + B_STAGE);
        consume_stage<A_KC, B_KC, TILE>(as, as + A_STAGE, acc, wm, wn, g, t, lane, empty0 + 8 * s);
    }

    if (SPLIT && csize > 1) {
        // Split-K hand-over through distributed shared memory: ranks > 0 park their partial accumulators in their own
        // (now idle) stage buffers, rank 0 pulls them over the cluster network in rank order and owns the epilogue.
        constexpr int NT = CONSUMER_WARPS * 32;
        __syncthreads();  // every DMMA warp of this CTA is done with the stage buffers
        if (!producer && crank != 0) {
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    tiles[((i * NJ + j) * 2 + 0) * NT + tid] = acc[i][j][0];
                    tiles[((i * NJ + j) * 2 + 1) * NT + tid] = acc[i][j][1];
                }
        }
        cluster_sync_all();
        if (!producer && crank == 0) {
            const uint32_t base = smem_u32(tiles);
            for (uint32_t r = 1; r < (uint32_t)csize; ++r) {
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        acc[i][j][0] += ld_cluster_f64(base + (uint32_t)((((i * NJ + j) * 2 + 0) * NT + tid) * 8), r);
                        acc[i][j][1] += ld_cluster_f64(base + (uint32_t)((((i * NJ + j) * 2 + 1) * NT + tid) * 8), r);
                    }
            }
        }
        cluster_sync_all();  // nobody's shared memory goes away before rank 0 has read it
        if (crank != 0) return;
    }
    if (producer) return;

    epilogue<TILE>(p, acc, m0, n0, wm, wn, g, t);
}

// ---- stream-K over 128 x 128 tiles ------------------------------------------------------------------------------------
// One 128 x 128 tile per CTA leaves the last wave partly empty: 2048^3 is 256 tiles on 148 SMs, two waves for 1.73 waves of
// work.  Here the k-iterations of the LAST (1 + remainder) waves of tiles are laid end to end (tile-major) and cut into
// sk_ctas equal ranges, one per CTA; the tiles before them stay one per CTA.  A range covers the tail of one tile, possibly
// whole tiles, and the head of another.  The CTA that holds a tile's FIRST k-slab owns its epilogue; a CTA whose range
// starts inside a tile computes that tail first, leaves the accumulators in its own 128 KB slot of the partial buffer and
// raises its flag - at the very start of its run, while the owner needs them only at the very end of its own, so in the
// steady state nobody waits.  The owner adds the partials after its own in ascending ticket (= ascending k) order, so the
// result depends on the shape and the SM count only, not on timing.
// Progress: CTAs take their position from a ticket.  A contributor never waits before it contributes, and tickets are
// handed out in start order, so the contributors of the oldest unfinished owner have either started or are next in line
// for a free SM: the grid finishes whenever two CTAs can be resident.
// The owner lowers each flag it consumed (every partial has exactly one consumer) and the last ticket zeroes the ticket
// counter: the buffer is all-zero again when the grid retires, so no epoch is needed and a captured graph can replay it.
struct StreamKArgs {
    double *partials;   // sk_ctas slots of 128 x 128 doubles, thread-major (slot[(frag * 256 + tid)])
    unsigned *flags;    // sk_ctas flags, zero between launches
    unsigned *ticket;   // zero between launches
    int sk_ctas;        // CTAs that share the stream-K region (tickets 0 .. sk_ctas-1)
    int sk_tiles;       // tiles of the stream-K region: linear tile ids 0 .. sk_tiles-1
    int ktiles;         // BK slabs per output tile
    unsigned grid;
};

__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(Shape<128>::THREADS, 1)
gemm_f64_tma_streamk(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TmaArgs p, StreamKArgs q)
{
    using S = Shape<128>;
    constexpr int MI = S::MI, NJ = S::NJ, CONSUMER_WARPS = S::CONSUMER_WARPS, NT = CONSUMER_WARPS * 32;
    constexpr int A_STAGE = S::A_STAGE, B_STAGE = S::B_STAGE;
    constexpr uint32_t STAGE_BYTES = S::STAGE_BYTES;
    static_assert(NT == 256, "consumers_sync() counts 256 threads");
    extern __shared__ uint8_t smem_raw[];
    __shared__ unsigned sm_ticket;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
    double *tiles = reinterpret_cast<double *>(smem_raw + pad);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + pad + (size_t)STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm_ticket = atomicAdd(q.ticket, 1u);
    }
    __syncthreads();
    const unsigned w = sm_ticket;
    const int KT = q.ktiles;
    const int64_t sk_iters = (int64_t)q.sk_tiles * KT;
    // this CTA's range [it0, it1) of the linear (tile, k-slab) iteration space
    int64_t it0, it1;
    if (w < (unsigned)q.sk_ctas) {
        it0 = sk_iters * w / q.sk_ctas;
        it1 = sk_iters * (w + 1) / q.sk_ctas;
    } else {
        it0 = ((int64_t)q.sk_tiles + (w - (unsigned)q.sk_ctas)) * KT;
        it1 = it0 + KT;
    }

    if (warp == CONSUMER_WARPS) {
        // ---- producer: the stage ring runs straight through the tile boundaries ----
        if (lane == 0) {
            int tile = (int)(it0 / KT);
            int gk = (int)(it0 - (int64_t)tile * KT);
            int tile_m, tile_n;
            raster(p, tile, tile_m, tile_n);
            const int total = (int)(it1 - it0);
            for (int kt = 0; kt < total; ++kt) {
                const int s = kt % STAGES;
                if (kt >= STAGES) mbar_wait(empty0 + 8 * s, (uint32_t)(kt / STAGES - 1) & 1u);
                const uint32_t a_dst = smem_u32(tiles + (size_t)s * (A_STAGE + B_STAGE));
                const uint32_t b_dst = a_dst + A_STAGE * sizeof(double);
                const int m0 = tile_m * 128, n0 = tile_n * 128;
                mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
                if constexpr (A_KC) tma_load_3d(a_dst, &map_a, full0 + 8 * s, 0, m0, gk * (BK / 4));
                else tma_load_3d(a_dst, &map_a, full0 + 8 * s, 0, gk * BK, m0 >> 2);
                if constexpr (B_KC) tma_load_3d(b_dst, &map_b, full0 + 8 * s, 0, n0, gk * (BK / 4));
                else tma_load_3d(b_dst, &map_b, full0 + 8 * s, 0, gk * BK, n0 >> 2);
                if (++gk == KT && kt + 1 < total) {
                    gk = 0;
                    raster(p, ++tile, tile_m, tile_n);
                }
            }
        }
        return;
    }

    // ---- consumers ----
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / S::WARPS_N, wn = warp % S::WARPS_N;
    double acc[MI][NJ][2];
    int kt = 0;
    for (int64_t it = it0; it < it1;) {
        const int tile = (int)(it / KT);
        const int kb = (int)(it - (int64_t)tile * KT);                 // first k-slab of this segment
        const int nseg = (int)min(it1 - it, (int64_t)(KT - kb));      // slabs of this tile that are mine
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int i = 0; i < nseg; ++i, ++kt) {
            const int s = kt % STAGES;
            mbar_wait(full0 + 8 * s, (uint32_t)(kt / STAGES) & 1u);
            const double *as = tiles + (size_t)s * (A_STAGE + B_STAGE);
            consume_stage<A_KC, B_KC, 128>(as, as + A_STAGE, acc, wm, wn, g, t, lane, empty0 + 8 * s);
        }
        it += nseg;
        if (kb == 0) {
            // the tile's first k-range is mine, so is its epilogue: collect what the CTAs after me computed for the rest
            // of it.  For them that was the FIRST thing they did, long ago by now: nobody waits in the steady state.
            if (nseg < KT) {
                const int64_t tile_end = (int64_t)(tile + 1) * KT;
                for (int c = (int)w + 1; c < q.sk_ctas && sk_iters * c / q.sk_ctas < tile_end; ++c) {
                    if (tid == 0) {
                        volatile unsigned *f = q.flags + c;
                        while (*f == 0u) {}
                        *f = 0u;  // consumed: every partial has exactly one reader
                        __threadfence();
                    }
                    consumers_sync();
                    const double *src = q.partials + (size_t)c * (128 * 128);
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            acc[i][j][0] += __ldcg(src + ((i * NJ + j) * 2 + 0) * NT + tid);
                            acc[i][j][1] += __ldcg(src + ((i * NJ + j) * 2 + 1) * NT + tid);
                        }
                }
            }
            int tile_m, tile_n;
            raster(p, tile, tile_m, tile_n);
            epilogue<128>(p, acc, (int64_t)tile_m * 128, (int64_t)tile_n * 128, wm, wn, g, t);
        } else {
            // my range starts inside this tile (only a CTA's first segment can): park the accumulators for the tile's owner
            double *dst = q.partials + (size_t)w * (128 * 128);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    __stcg(dst + ((i * NJ + j) * 2 + 0) * NT + tid, acc[i][j][0]);
                    __stcg(dst + ((i * NJ + j) * 2 + 1) * NT + tid, acc[i][j][1]);
                }
            __threadfence();
            consumers_sync();
            if (tid == 0) *(volatile unsigned *)(q.flags + w) = 1u;
        }
    }
    if (tid == 0 && w == q.grid - 1u) *(volatile unsigned *)q.ticket = 0u;  // every other CTA holds its ticket already
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x K operand with leading dimension ld; k_contig: element (r,k) at base[r*ld + k], else base[k*ld + r]
bool make_map(CUtensorMap *map, const double *base, int64_t rows, int64_t k, int64_t ld, bool k_contig, int tile)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    if (k_contig) {
        dims[0] = 4; dims[1] = (cuuint64_t)rows; dims[2] = (cuuint64_t)(k / 4);
        box[0] = 4; box[1] = (cuuint32_t)tile; box[2] = BK / 4;
    } else {
        dims[0] = 4; dims[1] = (cuuint64_t)k; dims[2] = (cuuint64_t)(rows / 4);
        box[0] = 4; box[1] = BK; box[2] = (cuuint32_t)tile / 4;
    }
    strides[0] = (cuuint64_t)ld * sizeof(double);
    strides[1] = 4 * sizeof(double);
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ticket + flags (zeroed once; the kernel leaves them zero) + one 128 KB partial slot per SM
int ensure_streamk(ndb_ctx *ctx, size_t bytes)
{
    if (ctx->streamk_bytes >= bytes) return NDB_OK;
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->streamk) NDB_CUDA(cudaFree(ctx->streamk));
    ctx->streamk = nullptr;
    ctx->streamk_bytes = 0;
    NDB_CUDA(cudaMalloc(&ctx->streamk, bytes));
    NDB_CUDA(cudaMemsetAsync(ctx->streamk, 0, bytes, ctx->stream));
    ctx->streamk_bytes = bytes;
    return NDB_OK;
}

size_t streamk_head(ndb_ctx *ctx) { return ((size_t)(64 + ctx->sm_count) * sizeof(unsigned) + 1023) & ~(size_t)1023; }
size_t streamk_bytes(ndb_ctx *ctx) { return streamk_head(ctx) + (size_t)ctx->sm_count * 128 * 128 * sizeof(double); }

int streamk_launch(ndb_ctx *ctx, const CUtensorMap &ma, const CUtensorMap &mb, const TmaArgs &p, bool a_kc, bool b_kc, int64_t out_tiles,
                   int64_t all_ktiles)
{
    const int sms = ctx->sm_count;
    const size_t head = streamk_head(ctx);
    NDB_TRY(ensure_streamk(ctx, streamk_bytes(ctx)));
    StreamKArgs q;
    q.ticket = reinterpret_cast<unsigned *>(ctx->streamk);
    q.flags = q.ticket + 64;
    q.partials = reinterpret_cast<double *>((char *)ctx->streamk + head);
    q.sk_ctas = sms;
    // the stream-K region is the last full wave plus the remainder (every CTA gets between one and two tiles of work),
    // or the whole problem when it is smaller than that
    q.sk_tiles = (int)(out_tiles > sms ? sms + out_tiles % sms : out_tiles);
    q.ktiles = (int)all_ktiles;
    q.grid = (unsigned)(q.sk_ctas + (out_tiles - q.sk_tiles));
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
#define NDB_F64_ATTR(A, B) \
    NDB_CUDA(cudaFuncSetAttribute(gemm_f64_tma_streamk<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape<128>::SMEM_BYTES));
        NDB_F64_ATTR(true, true) NDB_F64_ATTR(true, false) NDB_F64_ATTR(false, true) NDB_F64_ATTR(false, false)
#undef NDB_F64_ATTR
        attr_done.fetch_or(dev_bit, std::memory_order_release);
    }
    const dim3 grid(q.grid), block(Shape<128>::THREADS);
    const size_t smem = Shape<128>::SMEM_BYTES;
    if (a_kc && b_kc) gemm_f64_tma_streamk<true, true><<<grid, block, smem, ctx->stream>>>(ma, mb, p, q);
    else if (a_kc) gemm_f64_tma_streamk<true, false><<<grid, block, smem, ctx->stream>>>(ma, mb, p, q);
    else if (b_kc) gemm_f64_tma_streamk<false, true><<<grid, block, smem, ctx->stream>>>(ma, mb, p, q);
    else gemm_f64_tma_streamk<false, false><<<grid, block, smem, ctx->stream>>>(ma, mb, p, q);
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

}  // namespace

// Dense operands as classified by gemm_f64.cu: A is m x k (a_kc: k contiguous), B is k x n given as n "rows" of k
// (b_kc: k contiguous), C row-major with leading dimension ldc.  Returns NDB_ERR_UNSUPPORTED when the shape or the
// driver rules this path out (the caller then uses the cp.async kernel).
int gemm_f64_tma_launch(ndb_ctx *ctx, const double *a, int64_t lda, bool a_kc, const double *b, int64_t ldb, bool b_kc, double *c,
                        int64_t ldc, int64_t m, int64_t n, int64_t k, double alpha, double beta)
{
    if (getenv("NDB_GEMM_F64_CPASYNC")) return NDB_ERR_UNSUPPORTED;
    if (k < BK || m > INT32_MAX / 2 || n > INT32_MAX / 2 || k > INT32_MAX / 2) return NDB_ERR_UNSUPPORTED;
    if ((a_kc || b_kc) && (k & 3)) return NDB_ERR_UNSUPPORTED;
    if (!a_kc && (m & 3)) return NDB_ERR_UNSUPPORTED;
    if (!b_kc && (n & 3)) return NDB_ERR_UNSUPPORTED;
    if ((lda & 1) || (ldb & 1) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15)) return NDB_ERR_UNSUPPORTED;
    // 128 x 128 tiles unless they would cover less than three quarters of the SMs
    const int64_t big_tiles = ((m + 127) / 128) * ((n + 127) / 128);
    int tile = big_tiles * 4 < (int64_t)ctx->sm_count * 3 ? 64 : 128;
    if (getenv("NDB_GEMM_F64_TILE")) tile = atoi(getenv("NDB_GEMM_F64_TILE")) == 64 ? 64 : 128;
    CUtensorMap ma, mb;
    if (!make_map(&ma, a, m, k, lda, a_kc, tile) || !make_map(&mb, b, n, k, ldb, b_kc, tile)) return NDB_ERR_UNSUPPORTED;
    TmaArgs p;
    p.c = c; p.m = m; p.n = n; p.k = k; p.ldc = ldc;
    p.alpha = alpha; p.beta = beta;
    p.tiles_m = (int)((m + tile - 1) / tile);
    p.tiles_n = (int)((n + tile - 1) / tile);
    // small outputs: split K over a thread-block cluster (2/4/8 CTAs) until the grid covers ~2 CTAs per SM, keeping at
    // least two k-tiles per CTA
    const int64_t out_tiles = (int64_t)p.tiles_m * p.tiles_n;
    const int64_t all_ktiles = (k + BK - 1) / BK;
    int csize = 1;
    if (tile == 64)
        while (csize < 8 && out_tiles * csize * 2 <= (int64_t)ctx->sm_count * 2 && all_ktiles / (csize * 2) >= 2) csize *= 2;
    if (getenv("NDB_GEMM_F64_SPLITK")) {
        const int want = atoi(getenv("NDB_GEMM_F64_SPLITK"));
        csize = (want == 2 || want == 4 || want == 8) && all_ktiles >= want ? want : 1;
    }
    p.csize = csize;
    const unsigned blocks = (unsigned)(out_tiles * csize);
    // stream-K when whole-tile CTAs would idle more than 2 % of the machine in the last wave (or when asked to); measured
    // (benchmarks/gemm_streamk.py): 2048^3 555 -> 496 us, 1792 x 4096 x 1792 1097 -> 746 us, 3072^3 (3.89 waves) 1643 -> 1622 us
    const int64_t sms = ctx->sm_count;
    bool streamk = tile == 128 && out_tiles > sms && out_tiles % sms != 0 && all_ktiles >= 8 &&
                   ((out_tiles + sms - 1) / sms) * sms * 100 > out_tiles * 102;
    if (const char *e = getenv("NDB_GEMM_F64_STREAMK")) streamk = atoi(e) != 0 && tile == 128 && out_tiles * all_ktiles >= 2 * sms;
    // the hand-over buffer cannot be (re)allocated while the stream is being captured: whole tiles then
    if (streamk && ctx->streamk_bytes < streamk_bytes(ctx) && stream_is_capturing(ctx)) streamk = false;
    if (streamk) return streamk_launch(ctx, ma, mb, p, a_kc, b_kc, out_tiles, all_ktiles);
    // the opt-in is per DEVICE (a process may hold one context per GPU); setting it twice from racing threads is harmless
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
#define NDB_F64_ATTR(A, B, T) \
    NDB_CUDA(cudaFuncSetAttribute(gemm_f64_tma<A, B, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape<T>::SMEM_BYTES));
        NDB_F64_ATTR(true, true, 128) NDB_F64_ATTR(true, false, 128) NDB_F64_ATTR(false, true, 128) NDB_F64_ATTR(false, false, 128)
        NDB_F64_ATTR(true, true, 64) NDB_F64_ATTR(true, false, 64) NDB_F64_ATTR(false, true, 64) NDB_F64_ATTR(false, false, 64)
#undef NDB_F64_ATTR
        attr_done.fetch_or(dev_bit, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(blocks, 1, 1);
    cfg.stream = ctx->stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = csize > 1 ? 1 : 0;
#define NDB_F64_LAUNCH(A, B, T)                          \
    do {                                                 \
        cfg.blockDim = dim3(Shape<T>::THREADS, 1, 1);    \
        cfg.dynamicSmemBytes = Shape<T>::SMEM_BYTES;     \
        NDB_CUDA(cudaLaunchKernelEx(&cfg, gemm_f64_tma<A, B, T>, ma, mb, p)); \
    } while (0)
    if (tile == 128) {
        if (a_kc && b_kc) NDB_F64_LAUNCH(true, true, 128);
        else if (a_kc) NDB_F64_LAUNCH(true, false, 128);
        else if (b_kc) NDB_F64_LAUNCH(false, true, 128);
        else NDB_F64_LAUNCH(false, false, 128);
    } else {
        if (a_kc && b_kc) NDB_F64_LAUNCH(true, true, 64);
        else if (a_kc) NDB_F64_LAUNCH(true, false, 64);
        else if (b_kc) NDB_F64_LAUNCH(false, true, 64);
        else NDB_F64_LAUNCH(false, false, 64);
    }
#undef NDB_F64_LAUNCH
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

}  // namespace ndb
