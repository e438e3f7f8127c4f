/*
 * ndb200.h — C ABI of libndb200, the B200 (sm_100a) data-parallel core for ndarray.
 *
 * This is the drop-in boundary.  A Rust companion crate (or any FFI host) binds
 * exactly these symbols; nothing here mentions torch, C++ or CUDA types.
 *
 * Every view is described the way ndarray's `NdProducer` exposes it
 * (reference src/zip/ndproducer.rs:57-100, src/lib.rs:1304-1346):
 *   ptr    — DEVICE pointer to the *logical first element* (may be the highest
 *            address when strides are negative, src/lib.rs:1322-1328)
 *   dim    — axis lengths
 *   stride — SIGNED strides, counted in ELEMENTS (not bytes)
 *
 * Error convention: every entry point returns an `int` status.  0 = ok,
 * 1..6 mirror ndarray's `ErrorKind` (src/error.rs:43-57), the rest are ours.
 * Where ndarray panics, we return a status and the host wrapper raises with
 * ndarray's own message (`ndb_last_error()`).  Nothing aborts across the ABI.
 *
 * Threading: a context owns one device + one stream + a workspace.  Calls are
 * asynchronous on the context stream unless documented otherwise; a context
 * is not re-entrant, distinct contexts are independent.
 *
 * There is NO CPU fallback behind any of the compute entry points: if no CUDA
 * device is usable they return NDB_ERR_CUDA.
 */
#ifndef NDB200_H
#define NDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDB_MAX_NDIM 8 /* IxDyn views of higher rank are rejected with NDB_ERR_UNSUPPORTED */
#define NDB_MAX_PARTS 9 /* Zip takes 1..9 producers (reference src/zip/mod.rs:918-928, tests/azip.rs:481) */

/* ---- status codes ------------------------------------------------------ */
enum {
    NDB_OK = 0,
    NDB_ERR_INCOMPATIBLE_SHAPE = 1,  /* ErrorKind::IncompatibleShape  (src/error.rs:46) */
    NDB_ERR_INCOMPATIBLE_LAYOUT = 2, /* ErrorKind::IncompatibleLayout */
    NDB_ERR_RANGE_LIMITED = 3,       /* ErrorKind::RangeLimited */
    NDB_ERR_OUT_OF_BOUNDS = 4,       /* ErrorKind::OutOfBounds */
    NDB_ERR_UNSUPPORTED = 5,         /* ErrorKind::Unsupported */
    NDB_ERR_OVERFLOW = 6,            /* ErrorKind::Overflow */
    NDB_ERR_INVALID_ARG = 7,         /* ddof out of range etc. — ndarray panics (src/numeric/impl_numeric.rs:408) */
    NDB_ERR_EMPTY = 8,               /* result is `None` in ndarray (mean_axis of a 0-length axis, impl_numeric.rs:349) */
    NDB_ERR_DIV_BY_ZERO = 9,         /* integer / or % by zero, or MIN / -1: Rust panics */
    NDB_ERR_ALIASING = 10,           /* a written producer overlaps itself (dimension/mod.rs:58 dim_stride_overlap) */
    NDB_ERR_CUDA = 100               /* CUDA runtime failure; text in ndb_last_error() */
};

/* ---- element types ------------------------------------------------------ */
typedef enum {
    NDB_F32 = 0,
    NDB_F64 = 1,
    NDB_I32 = 2,
    NDB_I64 = 3
} ndb_dtype;

/* ---- a strided view (== what NdProducer exposes) ------------------------ */
typedef struct {
    void *ptr;                    /* device pointer to logical element [0,0,...] */
    int32_t ndim;                 /* 0..NDB_MAX_NDIM */
    int32_t dtype;                /* ndb_dtype */
    int64_t dim[NDB_MAX_NDIM];    /* axis lengths */
    int64_t stride[NDB_MAX_NDIM]; /* signed, in elements */
} ndb_view;

/* ---- elementwise op vocabulary ------------------------------------------
 * Closures cannot cross an FFI, so Zip::for_each / map_collect take an op from
 * this vocabulary: the closures the reference itself passes to Zip (operator
 * overloads src/impl_ops.rs:53-438,597-712; assign/fill src/impl_methods.rs:
 * 2796-2823; scaled_add src/linalg/impl_linalg.rs:399-405; the n-ary sums of
 * tests/azip.rs:481-497; config-1's a = b*c + a).
 * Parts are ordered OUTPUT FIRST: parts[0] is written, parts[1..] are read
 * (an input may be the same view as the output: `a += b` is ADD on (a; a, b)).
 * Float arithmetic is IEEE with every operation rounded separately (Rust never
 * contracts a*b+c); integers wrap; integer / and % follow Rust (truncating,
 * NDB_ERR_DIV_BY_ZERO where Rust would panic).
 */
typedef enum {
    NDB_OP_COPY = 0,     /* out = in0                     (assign, impl_methods.rs:2796) */
    NDB_OP_FILL = 1,     /* out = s0                      (fill,   impl_methods.rs:2819) */
    NDB_OP_ADD = 2,      /* out = in0 + in1 */
    NDB_OP_SUB = 3,      /* out = in0 - in1 */
    NDB_OP_MUL = 4,      /* out = in0 * in1 */
    NDB_OP_DIV = 5,      /* out = in0 / in1 */
    NDB_OP_REM = 6,      /* out = in0 % in1               (fmod for floats) */
    NDB_OP_MUL_ADD = 7,  /* out = in0 * in1 + in2         (two roundings; BASELINE config 1) */
    NDB_OP_ADD_S = 8,    /* out = in0 + s0                (array op scalar, impl_ops.rs:183) */
    NDB_OP_SUB_S = 9,    /* out = in0 - s0 */
    NDB_OP_MUL_S = 10,   /* out = in0 * s0 */
    NDB_OP_DIV_S = 11,   /* out = in0 / s0 */
    NDB_OP_REM_S = 12,   /* out = in0 % s0 */
    NDB_OP_RSUB_S = 13,  /* out = s0 - in0                (scalar op array, impl_ops.rs:348) */
    NDB_OP_RDIV_S = 14,  /* out = s0 / in0 */
    NDB_OP_NEG = 15,     /* out = -in0                    (impl_ops.rs Neg) */
    NDB_OP_AXPY = 16,    /* out = in0 + s0 * in1          (scaled_add, impl_linalg.rs:399) */
    NDB_OP_ADD_N = 17,   /* out = in0 + in1 + ... (left to right, 2..8 inputs; tests/azip.rs:492) */
    NDB_OP_ABS = 18,     /* out = |in0| */
    NDB_OP_SQRT = 19,    /* out = sqrt(in0)               (floats only; correctly rounded) */
    NDB_OP_MIN = 20,     /* out = min(in0, in1) */
    NDB_OP_MAX = 21,     /* out = max(in0, in1) */
    NDB_OP_SCALE_ADD = 22, /* out = in0 * s0 + in1 * s1   (gemv/gemm-style alpha/beta combine, two products one sum) */
    NDB_OP_TF32_HI = 23, /* out = tf32(in0), round to nearest (f32 only; operand packing of the 3xTF32 GEMM) */
    NDB_OP_TF32_LO = 24, /* out = in0 - tf32(in0)         (f32 only) */
    /* predicates (closures given to Zip::all / Zip::any, src/zip/mod.rs:658-696): out = 1 or 0 in the element type */
    NDB_OP_EQ = 25,
    NDB_OP_NE = 26,
    NDB_OP_LT = 27,
    NDB_OP_LE = 28,
    NDB_OP_GT = 29,
    NDB_OP_GE = 30,
    /* float maps of src/numeric/impl_float_maths.rs:80-160 (mapv with a libm function; floats only).  The function
     * is a RUN-TIME selector so one kernel serves all of them: s0 = (ndb_math_fn) as the element type, s1 = parameter
     * (powi exponent, powf exponent, log base, abs_sub / hypot other operand).  Exactly-rounded functions (floor,
     * ceil, round, trunc, fract, signum, recip, pow2, is_nan, is_infinite) are bit-exact; transcendental ones are
     * CUDA libm (<= 2 ulp) and compared to the CPU within a tolerance. */
    NDB_OP_MATH = 31,
    NDB_OP_CLAMP = 32, /* out = clamp(in0, s0, s1)  (impl_float_maths.rs clamp; NaN stays NaN) */
    NDB_OP__COUNT
} ndb_op;

typedef enum {
    NDB_FN_FLOOR = 0, NDB_FN_CEIL, NDB_FN_ROUND, NDB_FN_TRUNC, NDB_FN_FRACT, NDB_FN_SIGNUM, NDB_FN_RECIP, NDB_FN_EXP,
    NDB_FN_EXP2, NDB_FN_EXP_M1, NDB_FN_LN, NDB_FN_LOG2, NDB_FN_LOG10, NDB_FN_LN_1P, NDB_FN_CBRT, NDB_FN_SIN, NDB_FN_COS,
    NDB_FN_TAN, NDB_FN_ASIN, NDB_FN_ACOS, NDB_FN_ATAN, NDB_FN_SINH, NDB_FN_COSH, NDB_FN_TANH, NDB_FN_ASINH, NDB_FN_ACOSH,
    NDB_FN_ATANH, NDB_FN_TO_DEGREES, NDB_FN_TO_RADIANS, NDB_FN_POW2, NDB_FN_POWI, NDB_FN_POWF, NDB_FN_LOG, NDB_FN_ABS_SUB,
    NDB_FN_HYPOT, NDB_FN_IS_NAN, NDB_FN_IS_INFINITE, NDB_FN__COUNT
} ndb_math_fn;

/* ---- axis-reduction vocabulary -------------------------------------------
 * sum_axis / product_axis / mean_axis / var_axis / std_axis
 * (src/numeric/impl_numeric.rs:268,305,342,401,471) and the associative folds
 * fold_axis is used with (src/impl_methods.rs:3130).
 */
typedef enum {
    NDB_RED_SUM = 0,
    NDB_RED_PROD = 1,
    NDB_RED_MIN = 2,
    NDB_RED_MAX = 3,
    NDB_RED_MEAN = 4, /* sum / n in the element type (integer division truncates) */
    NDB_RED_VAR = 5,  /* Welford, result M2 / (n - ddof)            (floats only) */
    NDB_RED_STD = 6,  /* sqrt(var)                                   (floats only) */
    NDB_RED__COUNT
} ndb_redop;

/* kernel classes the planner can choose (reported by ndb_zip_plan) */
typedef enum {
    NDB_KERNEL_NOOP = 0,      /* some axis has length 0 */
    NDB_KERNEL_FLAT_VEC = 1,  /* every producer is one dense unit-stride run: 128-bit vector kernel */
    NDB_KERNEL_FLAT = 2,      /* same, but a base pointer is not 16-byte aligned: scalar coalesced */
    NDB_KERNEL_ROWS_VEC = 3,  /* collapsed N-d, unit (or 0) inner strides, aligned rows: vector kernel */
    NDB_KERNEL_ROWS = 4,      /* collapsed N-d, arbitrary signed strides, coalesced along the output's fastest axis */
    NDB_KERNEL_TILE = 5       /* producers disagree on the fastest axis: shared-memory tile transpose */
} ndb_kernel_class;

typedef struct {
    int32_t kernel;               /* ndb_kernel_class */
    int32_t ndim;                 /* collapsed rank */
    int64_t dim[NDB_MAX_NDIM];    /* collapsed axis lengths, slowest first */
    int64_t stride[NDB_MAX_PARTS][NDB_MAX_NDIM]; /* per producer collapsed strides */
    int64_t offset[NDB_MAX_PARTS]; /* element offset applied to each producer's ptr (axis flips) */
    int32_t tile_axis;            /* NDB_KERNEL_TILE: collapsed axis staged through shared memory, else -1 */
    uint32_t staged_mask;         /* NDB_KERNEL_TILE: bit p set = producer p is transposed through shared memory */
} ndb_zip_plan_t;

typedef struct ndb_ctx ndb_ctx;

/* ---- library / context --------------------------------------------------- */
const char *ndb_version(void);
/* Thread-local text of the last failure on this thread ("" if none). */
const char *ndb_last_error(void);
/* Number of kernels this library has launched in the calling process (all contexts). */
uint64_t ndb_launch_count(void);
int ndb_device_count(int *count);

int ndb_ctx_create(int device, ndb_ctx **out);
int ndb_ctx_destroy(ndb_ctx *ctx);
/* Use a caller-owned cudaStream_t (passed as void*) instead of the context's own. NULL restores it. */
int ndb_ctx_set_stream(ndb_ctx *ctx, void *cuda_stream);
void *ndb_ctx_stream(ndb_ctx *ctx);
int ndb_sync(ndb_ctx *ctx);

/* ---- device memory (DeviceArray storage; analogue of OwnedRepr, src/data_repr.rs:24) */
int ndb_alloc(ndb_ctx *ctx, size_t bytes, void **dptr);
int ndb_free(ndb_ctx *ctx, void *dptr);
int ndb_memset(ndb_ctx *ctx, void *dptr, int byte, size_t bytes);
int ndb_h2d(ndb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async on ctx stream */
int ndb_d2h(ndb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async on ctx stream */
int ndb_d2d(ndb_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes);
int ndb_host_alloc(size_t bytes, void **hptr); /* pinned host memory */
int ndb_host_free(void *hptr);
/* event timing on the context stream, for harnesses */
int ndb_event_create(void **ev);
int ndb_event_destroy(void *ev);
int ndb_event_record(ndb_ctx *ctx, void *ev);
int ndb_event_elapsed_ms(void *start, void *stop, float *ms); /* synchronises on `stop` */

/* CUDA graphs: record the calls issued on the context between begin and end (asynchronous entry points only: Zip, axis
 * reductions, scans, GEMM - nothing that returns a host scalar or grows a workspace for the first time) and replay them
 * with one launch.  Kernels keep their programmatic-dependent-launch edges inside the graph.  For loops of ~10 us kernels,
 * where the host cannot issue launches as fast as the GPU retires them. */
int ndb_graph_begin(ndb_ctx *ctx);
int ndb_graph_end(ndb_ctx *ctx, void **graph_out);
int ndb_graph_launch(ndb_ctx *ctx, void *graph);
int ndb_graph_destroy(void *graph);

/* ---- view arithmetic: pure host code, identical to ndarray's ---------------
 * All of these only rewrite (ptr, dim, stride); `elem_size` comes from dtype.
 */
/* Slice one axis with start/end/step exactly like `Slice{start,end,step}`
 * (src/slice.rs:42, src/dimension/mod.rs:390-495 do_slice).  `has_end == 0` means `end = None`.
 * Out-of-range → NDB_ERR_OUT_OF_BOUNDS, step == 0 → NDB_ERR_INVALID_ARG (both panic in ndarray). */
int ndb_view_slice_axis(ndb_view *v, int axis, int64_t start, int64_t end, int has_end, int64_t step);
/* index_axis / collapse: fix `axis` at `index` and remove it (src/impl_methods.rs:1009). */
int ndb_view_index_axis(ndb_view *v, int axis, int64_t index);
/* broadcast to `ndim`/`dim` (src/impl_methods.rs:2371-2430 upcast): NDB_ERR_INCOMPATIBLE_SHAPE if impossible (ndarray returns None). */
int ndb_view_broadcast(const ndb_view *v, int ndim, const int64_t *dim, ndb_view *out);
/* permuted_axes (src/impl_methods.rs:2522): out axis i = in axis perm[i]. */
int ndb_view_permute(const ndb_view *v, const int *perm, ndb_view *out);
/* invert_axis (src/impl_methods.rs:2676). */
int ndb_view_invert_axis(ndb_view *v, int axis);
/* merge_axes (src/dimension/mod.rs:739-763, src/impl_methods.rs:2724): fold axis `take` into axis `into` if the
 * memory layout allows; returns 1 if merged (view updated), 0 if not possible (view unchanged), <0 on bad axes. */
int ndb_view_merge_axes(ndb_view *v, int take, int into);
/* collapse_axis (src/impl_methods.rs:1075): keep only `index` along `axis`; the axis stays with length 1. */
int ndb_view_collapse_axis(ndb_view *v, int axis, int64_t index);
/* split_at (src/impl_raw_views.rs:113-135): the sharding primitive. */
int ndb_view_split_at(const ndb_view *v, int axis, int64_t index, ndb_view *left, ndb_view *right);
/* Layout bitset of src/layout/mod.rs:9-16 as computed by array_layout (src/zip/mod.rs:54-77):
 * bit0 CORDER, bit1 FORDER, bit2 CPREFER, bit3 FPREFER. */
uint32_t ndb_view_layout(const ndb_view *v);
/* 1 if the view is one dense block in some axis order with any stride signs (dimension_trait.rs:294). */
int ndb_view_is_contiguous(const ndb_view *v);
/* min_stride_axis (dimension_trait.rs:343; Ix2 special case :715) — decides sum_axis' fast path. */
int ndb_view_min_stride_axis(const ndb_view *v);

/* ---- Zip: n-ary lock-step elementwise traversal -----------------------------
 * Replaces Zip::from(..).and(..).for_each(closure) (src/zip/mod.rs:212,704,585,309-424)
 * and its rayon form par_for_each (src/parallel/impl_par_methods.rs:64).
 * parts[0] is the output producer, parts[1..n_parts) the inputs; all shapes must be equal
 * (NDB_ERR_INCOMPATIBLE_SHAPE — ndarray panics "Zip: Producer dimension mismatch", zip/mod.rs:252).
 * `scalars` points at 0..2 values of the element type (FILL, *_S, AXPY, SCALE_ADD).
 */
int ndb_zip(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts);
/* out[i] = in[i] as <out dtype> — `mapv(|x| x as B)`: the one elementwise form with two element types.  Rust `as`
 * semantics (float -> int truncates and saturates, NaN -> 0; int -> int wraps; f64 -> f32 rounds to nearest). */
int ndb_cast(ndb_ctx *ctx, const ndb_view *out, const ndb_view *in);
/* The plan ndb_zip would execute; pure host code (no device needed). */
int ndb_zip_plan(int op, int n_parts, const ndb_view *parts, ndb_zip_plan_t *plan);
/* Zip::all / Zip::any (src/zip/mod.rs:658-696): does the predicate `op` (NDB_OP_EQ..GE, or any op whose result is
 * compared with zero) hold for every / some element tuple?  `inputs` are the producers (none is written); an empty
 * shape gives all = 1 (vacuously true, tests/azip.rs:472-478) and any = 0.  Synchronises. */
int ndb_zip_all(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, int *result);
int ndb_zip_any(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, int *result);
/* Zip::fold / par_fold with an additive accumulator (src/zip/mod.rs:622, parallel/impl_par_methods.rs:173):
 * sum over all element tuples of op(inputs...), result (element type) to host. */
int ndb_zip_fold_sum(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, void *out_host);
/* Zip::fold with any ASSOCIATIVE accumulator (src/zip/mod.rs:622-655): out = redop(init, redop over all element tuples of
 * op(inputs...)), redop in {SUM, PROD, MIN, MAX}; `init` (element type, host) may be NULL = the identity.  An empty shape
 * returns `init` like Rust's fold (NDB_ERR_EMPTY for MIN/MAX without one).  Closures that are not associative cannot be
 * evaluated as a tree and are out of scope (fold_axis has the same restriction). */
int ndb_zip_fold(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, int redop, const void *init,
                 void *out_host);
/* Zip::fold whose closure also WRITES through its first producer, e.g.
 *     Zip::from(&mut a).and(&b).and(&c).fold(0., |acc, a, &b, &c| { *a = b * c + *a; acc + *a })
 * parts are as in ndb_zip (parts[0] written, the rest read), and the values written are folded with redop into the DEVICE
 * scalar `out_dev` (element type; asynchronous on the context stream).  Dense producers take one fused pass (elementwise +
 * reduction: BASELINE config 5 at 16 bytes per element instead of 20); anything else runs the Zip and then the reduction.
 * across_ranks != 0: the kernel that finishes the fold also exchanges the scalar through the peer windows, every rank
 * receives the fold over ALL shards (windows must be attached). */
int ndb_zip_fold_write(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts, int redop, void *out_dev,
                       int across_ranks);
/* Zip::fold_while (src/zip/mod.rs:640-655) for the same associative accumulators.  FoldWhile::Done only means something
 * relative to a traversal order, so the order is kept explicit: the values op(inputs...) are laid out in Zip's traversal
 * order (C, or F when the producers prefer it, zip/mod.rs:278-281), the FIRST position whose value satisfies
 * `value <stop_op> *stop_value` (stop_op in NDB_OP_EQ..NDB_OP_GE) is found, and the fold covers the elements before it:
 *     fold_while(init, |acc, ..| if v <stop_op> s { Done(acc) } else { Continue(acc <redop> v) })
 * *done = 1 if the predicate fired (FoldWhile::Done), 0 if every element was folded (Continue). */
int ndb_zip_fold_while(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, int redop, const void *init,
                       int stop_op, const void *stop_value, void *out_host, int *done);
/* Zip::map_collect's output-order rule (src/zip/mod.rs:278-281,764-789): returns 1 if the fresh
 * output must be F-order, 0 for C-order, given the INPUT producers only. */
int ndb_zip_prefer_f(int n_inputs, const ndb_view *inputs);

/* ---- reductions ------------------------------------------------------------- */
/* ArrayRef::sum / product (src/numeric/impl_numeric.rs:32,83): reduce every element, result
 * written to HOST memory `out` (element type); synchronises the context stream. op ∈ {SUM, PROD, MIN, MAX}. */
int ndb_reduce_all(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_host);
/* Same, but the scalar stays on the device (`out_dev`, element type) and nothing synchronises:
 * used by the multi-GPU path so the NCCL allreduce can follow on the same stream. */
int ndb_reduce_all_dev(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_dev);
/* sum_axis / product_axis / mean_axis / var_axis / std_axis / fold_axis(min|max|sum|prod).
 * `out` must have `in`'s shape with `axis` removed (any strides).  `ddof` is used by VAR/STD only
 * and must satisfy 0 <= ddof <= n (NDB_ERR_INVALID_ARG otherwise; NaN passes, as in ndarray). */
int ndb_reduce_axis(ndb_ctx *ctx, int redop, int axis, const ndb_view *in, const ndb_view *out, double ddof);
/* Partial Welford state for the multi-GPU var_axis: writes mean and M2 (sum of squared deviations)
 * of this shard along `axis` into `mean_out` / `m2_out`, and merges shards with Chan's formula. */
int ndb_welford_axis(ndb_ctx *ctx, int axis, const ndb_view *in, const ndb_view *mean_out, const ndb_view *m2_out);
int ndb_welford_merge(ndb_ctx *ctx, int n_shards, const int64_t *counts, const ndb_view *means, const ndb_view *m2s,
                      double ddof, int want_std, const ndb_view *out);
/* ArrayRef::var / std over every element (src/numeric/impl_numeric.rs:180-249), result to host (floats only). */
int ndb_var(ndb_ctx *ctx, const ndb_view *in, double ddof, int want_std, void *out_host);
/* ---- multi-GPU exchange over peer memory (NVLink), one process per GPU -------------------------------------------
 * The sharded reductions of SURVEY.md 8e end in a small exchange: sum()/dot -> one element, sum_axis(Axis(0)) -> the
 * [n_out] vector, var_axis(Axis(0)) -> per-shard (mean, M2).  Each rank creates a window in its own HBM and publishes its
 * 64-byte CUDA IPC handle; after every rank has attached all handles, an exchange is ONE kernel launch per rank that
 * stores this rank's contribution into every rank's window, signals, waits for the others and combines in rank order
 * (deterministic).  All ranks must issue the same sequence of exchanges with the same sizes.  slot_bytes bounds the size
 * of one contribution.  (NCCL through torch.distributed remains available in ndarray_b200/parallel.py as the fallback.) */
int ndb_peer_window_create(ndb_ctx *ctx, int rank, int world, size_t slot_bytes, void *handle_out /* 64 bytes */);
int ndb_peer_window_attach(ndb_ctx *ctx, const void *handles /* world x 64 bytes, in rank order */);
int ndb_peer_window_destroy(ndb_ctx *ctx);
/* out[i] = redop over ranks (rank order) of local[i]; redop in {SUM, PROD, MIN, MAX}; out may alias local. */
int ndb_peer_allreduce(ndb_ctx *ctx, int redop, int dtype, const void *local_dev, void *out_dev, int64_t n);
/* out[r*n + i] = rank r's local[i]. */
int ndb_peer_allgather(ndb_ctx *ctx, int dtype, const void *local_dev, void *out_dev, int64_t n);
/* ndb_reduce_axis over ALL shards when `axis` is the sharded axis and a kept axis is the dense one (sum_axis(Axis(0)) of a
 * C-order slab): ONE launch per rank - the block that completes a column tile locally exchanges its states (sums, or
 * Welford (n, mean, M2) for var/std) with the other ranks through the windows, merges them in rank order and finalizes
 * with the GLOBAL axis length n_global (mean divides by it, ddof is checked against it).
 * Contract: every rank passes its slab of ONE global array - same kept shape, strides, dtype and op; the row count may
 * differ and may be 0.  Returns NDB_ERR_UNSUPPORTED, without launching anything, when the layout, the tile count or the slot
 * size rule this path out.  That rule reads only values the contract makes identical on every rank (kept shape, strides,
 * dtype, op, n_global, slot size, occupancy x SM count): a shard that starts off a 16-byte boundary or holds a single row
 * changes how the rank loads its rows, not whether it takes part, so all ranks decide alike and the caller falls back to
 * ndb_reduce_axis + ndb_peer_allreduce together.  The launch is cooperative (every block co-resident or the launch fails),
 * so a block waiting for its peers can never keep the blocks it waits for off the machine. */
int ndb_reduce_axis_peer(ndb_ctx *ctx, int redop, int axis, const ndb_view *in, const ndb_view *out, double ddof, int64_t n_global);
/* Would ndb_reduce_axis_peer take the fused path for these arguments?  NDB_OK / NDB_ERR_UNSUPPORTED, launches nothing; the
 * answer is the same on every rank under the contract above. */
int ndb_reduce_axis_peer_ok(ndb_ctx *ctx, int redop, int axis, const ndb_view *in, const ndb_view *out, double ddof, int64_t n_global);
/* ndb_reduce_all_dev over ALL shards: the kernel that finishes this rank's reduction tree also performs the exchange. */
int ndb_reduce_all_peer(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_dev);

/* ---- sharded entry points: everything a non-Python host needs to run the multi-GPU paths of SURVEY.md 8e -------------------
 * One process per GPU, Axis(0) is the sharded axis.  The communicator is NCCL's, owned by the library: libnccl.so.2 is
 * dlopen'ed on first use (NDB_NCCL_LIB overrides the name), so libndb200 has no link-time NCCL dependency.  Rank 0 calls
 * ndb_comm_unique_id and ships the 128 bytes to the other ranks by any means (file, socket, MPI, torch.distributed); every
 * rank then calls ndb_comm_init.  Collectives run on the context stream; all ranks must issue the same sequence of calls. */
int ndb_comm_unique_id(void *id_out /* 128 bytes */);
int ndb_comm_init(ndb_ctx *ctx, int rank, int world, const void *id /* 128 bytes */);
int ndb_comm_destroy(ndb_ctx *ctx);
/* Replicate a dense device array from rank `src` (ncclBroadcast on the context stream). */
int ndb_broadcast(ndb_ctx *ctx, const ndb_view *v, int src);
/* general_mat_mul with A and C sharded along Axis(0) and B replicated: `b` is a dense [k, n] buffer on every rank whose
 * contents are valid on rank `src`; it is broadcast, then every rank multiplies its own rows (row sharding changes no
 * row's arithmetic).  With one rank (no communicator) this is ndb_gemm. */
int ndb_gemm_rowsharded(ndb_ctx *ctx, const void *alpha, const ndb_view *a_local, const ndb_view *b, const void *beta,
                        const ndb_view *c_local, int src);
/* sum_axis / product_axis / mean_axis / var_axis / std_axis / fold_axis(min|max) over ALL shards.  axis != 0: plain
 * ndb_reduce_axis (outputs are sharded like the input).  axis == 0: the fused cooperative kernel over the peer windows when
 * they are attached and the layout qualifies (ndb_reduce_axis_peer), else local partial + all-reduce (mean divides by
 * n_global afterwards), or for var/std per-shard Welford states + all-gather + Chan merge in rank order (deterministic).
 * Every rank receives the full result in `out`.  A rank whose shard is EMPTY (fewer rows than ranks) still takes part and
 * must make the call: it contributes the identity (min / max: the element type's extreme); only n_global == 0 - the same
 * on every rank - gives NDB_ERR_EMPTY for mean / min / max, before anybody enters the exchange. */
int ndb_reduce_axis_sharded(ndb_ctx *ctx, int redop, int axis, const ndb_view *in, const ndb_view *out, double ddof, int64_t n_global);
/* sum() / product() / min / max over all shards, result (element type) to host on every rank: the kernel that finishes the
 * local tree also does the exchange when peer windows are attached, else an NCCL all-reduce of one element follows.
 * An empty shard contributes the identity (min / max: the type's extreme, which is also the result when EVERY shard is
 * empty - the entry point does not know the global length). */
int ndb_reduce_all_sharded(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_host);

/* accumulate_axis_inplace (src/impl_methods.rs:3243) with op NDB_OP_ADD (running sum) or NDB_OP_MUL (cumprod,
 * src/numeric/impl_numeric.rs:125): v[.., i, ..] = op(v[.., i-1, ..], v[.., i, ..]) along `axis`, in place, one launch. */
int ndb_accumulate_axis(ndb_ctx *ctx, int op, int axis, const ndb_view *v);
/* 1-D dot product (src/linalg/impl_linalg.rs:74-133, numeric_util.rs:56), result to host. */
int ndb_dot(ndb_ctx *ctx, const ndb_view *x, const ndb_view *y, void *out_host);

/* ---- linalg ------------------------------------------------------------------
 * general_mat_mul(alpha, &a, &b, beta, &mut c) (src/linalg/impl_linalg.rs:633): C <- alpha*A*B + beta*C,
 * any signed strides like matrixmultiply::{s,d}gemm (impl_linalg.rs:512-546). alpha/beta point at HOST
 * scalars of the element type. With beta == 0, C is never read (impl_linalg.rs:323-329).
 * f64 uses FP64 tensor-core DMMA, f32 uses 3xTF32 on tcgen05 with fp32-class accuracy. */
int ndb_gemm(ndb_ctx *ctx, const void *alpha, const ndb_view *a, const ndb_view *b, const void *beta, const ndb_view *c);
/* general_mat_vec_mul (impl_linalg.rs:657): y <- alpha*A*x + beta*y */
int ndb_gemv(ndb_ctx *ctx, const void *alpha, const ndb_view *a, const ndb_view *x, const void *beta, const ndb_view *y);

/* ---- CBLAS drop-in: ndarray's one real plugin boundary -------------------------
 * ndarray built with feature "blas" calls these with HOST pointers
 * (src/linalg/impl_linalg.rs:466-491 gemm, :116-122 dot); crates/blas-mock-tests/src/lib.rs
 * shows a foreign `extern "C" cblas_dgemm` is a legitimate provider.  These stage H2D/D2H around
 * ndb_gemm on a process-global context (device NDB_DEVICE or 0).
 */
enum { NdbCblasRowMajor = 101, NdbCblasColMajor = 102 };
enum { NdbCblasNoTrans = 111, NdbCblasTrans = 112, NdbCblasConjTrans = 113 };
void cblas_sgemm(int layout, int transa, int transb, int m, int n, int k, float alpha, const float *a, int lda,
                 const float *b, int ldb, float beta, float *c, int ldc);
void cblas_dgemm(int layout, int transa, int transb, int m, int n, int k, double alpha, const double *a, int lda,
                 const double *b, int ldb, double beta, double *c, int ldc);
void cblas_sgemv(int layout, int trans, int m, int n, float alpha, const float *a, int lda, const float *x, int incx,
                 float beta, float *y, int incy);
void cblas_dgemv(int layout, int trans, int m, int n, double alpha, const double *a, int lda, const double *x, int incx,
                 double beta, double *y, int incy);
float cblas_sdot(int n, const float *x, int incx, const float *y, int incy);
double cblas_ddot(int n, const double *x, int incx, const double *y, int incy);

/* ---- host-buffer convenience (end-to-end path) ----------------------------------
 * Same semantics as ndb_zip but every view's `ptr` is a HOST pointer (pinned for full speed);
 * inputs are copied H2D, the op runs, the output is copied D2H, chunked along the outermost
 * axis so copies in both directions overlap the kernel.  Synchronous. */
int ndb_zip_host(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts);
/* The same Zip followed by a fold of the values it wrote (Zip::fold with an associative accumulator, src/zip/mod.rs:622):
 * each slab is reduced on the device while it is still resident, the per-slab scalars are combined in slab order and
 * the result (element type) lands in `out_host`.  redop in {SUM, PROD, MIN, MAX}.  BASELINE config 5 end to end is
 * ndb_zip_host_fold(MUL_ADD, (a; b, c, a), SUM): a = b*c + a and a.sum() in one pass over PCIe. */
int ndb_zip_host_fold(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts, int redop, void *out_host);

/* ---- synthetic data for harnesses ----------------------------------------------------------------------------------
 * Counter-based generator: element i of the stream is a pure function of (seed, i) (splitmix64 output i, top bits mapped to
 * [-1, 1) for floats, raw bits for integers), so shards of one logical array can be filled independently on any number
 * of GPUs and reproduced bit for bit on the CPU (oracle: orc_counter_fill).  `out` must be standard-layout and 16-byte
 * aligned; out[j] (C-order flat index j) = value(seed, first_index + j).  Not part of ndarray's path. */
int ndb_fill_counter_rng(ndb_ctx *ctx, const ndb_view *out, uint64_t seed, uint64_t first_index);

#ifdef __cplusplus
}
#endif
#endif /* NDB200_H */
