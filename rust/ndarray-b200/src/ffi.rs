//! Raw bindings of include/ndb200.h.  One `extern "C"` item per declaration, nothing else.
#![allow(non_camel_case_types)]
use libc::{c_char, c_double, c_int, c_void, size_t};

pub const NDB_MAX_NDIM: usize = 8;
pub const NDB_MAX_PARTS: usize = 9;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ndb_view {
    pub ptr: *mut c_void,
    pub ndim: i32,
    pub dtype: i32,
    pub dim: [i64; NDB_MAX_NDIM],
    pub stride: [i64; NDB_MAX_NDIM],
}

#[repr(C)]
pub struct ndb_ctx { _private: [u8; 0] }

/// mirrors `ndb_zip_plan_t` (include/ndb200.h): what the planner decided for a Zip
#[repr(C)]
pub struct ndb_zip_plan_t {
    pub kernel: i32,
    pub ndim: i32,
    pub dim: [i64; NDB_MAX_NDIM],
    pub stride: [[i64; NDB_MAX_NDIM]; NDB_MAX_PARTS],
    pub offset: [i64; NDB_MAX_PARTS],
    pub tile_axis: i32,
    pub staged_mask: u32,
}

pub const NDB_OK: c_int = 0;
pub const NDB_ERR_INCOMPATIBLE_SHAPE: c_int = 1;
pub const NDB_ERR_EMPTY: c_int = 8;

extern "C" {
    pub fn ndb_version() -> *const c_char;
    pub fn ndb_last_error() -> *const c_char;
    pub fn ndb_ctx_create(device: c_int, out: *mut *mut ndb_ctx) -> c_int;
    pub fn ndb_ctx_destroy(ctx: *mut ndb_ctx) -> c_int;
    pub fn ndb_sync(ctx: *mut ndb_ctx) -> c_int;
    pub fn ndb_alloc(ctx: *mut ndb_ctx, bytes: size_t, dptr: *mut *mut c_void) -> c_int;
    pub fn ndb_free(ctx: *mut ndb_ctx, dptr: *mut c_void) -> c_int;
    pub fn ndb_memset(ctx: *mut ndb_ctx, dptr: *mut c_void, byte: c_int, bytes: size_t) -> c_int;
    pub fn ndb_h2d(ctx: *mut ndb_ctx, dst: *mut c_void, src: *const c_void, bytes: size_t) -> c_int;
    pub fn ndb_d2h(ctx: *mut ndb_ctx, dst: *mut c_void, src: *const c_void, bytes: size_t) -> c_int;
    pub fn ndb_view_slice_axis(v: *mut ndb_view, axis: c_int, start: i64, end: i64, has_end: c_int, step: i64) -> c_int;
    pub fn ndb_view_broadcast(v: *const ndb_view, ndim: c_int, dim: *const i64, out: *mut ndb_view) -> c_int;
    pub fn ndb_view_permute(v: *const ndb_view, perm: *const c_int, out: *mut ndb_view) -> c_int;
    pub fn ndb_view_invert_axis(v: *mut ndb_view, axis: c_int) -> c_int;
    pub fn ndb_view_split_at(v: *const ndb_view, axis: c_int, index: i64, l: *mut ndb_view, r: *mut ndb_view) -> c_int;
    pub fn ndb_zip(ctx: *mut ndb_ctx, op: c_int, scalars: *const c_void, n_parts: c_int, parts: *const ndb_view) -> c_int;
    pub fn ndb_zip_prefer_f(n_inputs: c_int, inputs: *const ndb_view) -> c_int;
    pub fn ndb_reduce_all(ctx: *mut ndb_ctx, redop: c_int, input: *const ndb_view, out_host: *mut c_void) -> c_int;
    pub fn ndb_reduce_axis(ctx: *mut ndb_ctx, redop: c_int, axis: c_int, input: *const ndb_view, out: *const ndb_view, ddof: c_double) -> c_int;
    pub fn ndb_dot(ctx: *mut ndb_ctx, x: *const ndb_view, y: *const ndb_view, out_host: *mut c_void) -> c_int;
    pub fn ndb_gemm(ctx: *mut ndb_ctx, alpha: *const c_void, a: *const ndb_view, b: *const ndb_view, beta: *const c_void, c: *const ndb_view) -> c_int;
    pub fn ndb_gemv(ctx: *mut ndb_ctx, alpha: *const c_void, a: *const ndb_view, x: *const ndb_view, beta: *const c_void, y: *const ndb_view) -> c_int;
    pub fn ndb_launch_count() -> u64;
    pub fn ndb_device_count(count: *mut c_int) -> c_int;
    pub fn ndb_ctx_set_stream(ctx: *mut ndb_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn ndb_ctx_stream(ctx: *mut ndb_ctx) -> *mut c_void;
    pub fn ndb_d2d(ctx: *mut ndb_ctx, dst: *mut c_void, src: *const c_void, bytes: size_t) -> c_int;
    pub fn ndb_host_alloc(bytes: size_t, hptr: *mut *mut c_void) -> c_int;
    pub fn ndb_host_free(hptr: *mut c_void) -> c_int;
    pub fn ndb_event_create(ev: *mut *mut c_void) -> c_int;
    pub fn ndb_event_destroy(ev: *mut c_void) -> c_int;
    pub fn ndb_event_record(ctx: *mut ndb_ctx, ev: *mut c_void) -> c_int;
    pub fn ndb_event_elapsed_ms(start: *mut c_void, stop: *mut c_void, ms: *mut f32) -> c_int;
    pub fn ndb_view_index_axis(v: *mut ndb_view, axis: c_int, index: i64) -> c_int;
    pub fn ndb_view_merge_axes(v: *mut ndb_view, take: c_int, into: c_int) -> c_int;
    pub fn ndb_view_collapse_axis(v: *mut ndb_view, axis: c_int, index: i64) -> c_int;
    pub fn ndb_view_layout(v: *const ndb_view) -> u32;
    pub fn ndb_view_is_contiguous(v: *const ndb_view) -> c_int;
    pub fn ndb_view_min_stride_axis(v: *const ndb_view) -> c_int;
    pub fn ndb_cast(ctx: *mut ndb_ctx, out: *const ndb_view, input: *const ndb_view) -> c_int;
    pub fn ndb_zip_plan(op: c_int, n_parts: c_int, parts: *const ndb_view, plan: *mut ndb_zip_plan_t) -> c_int;
    pub fn ndb_zip_all(ctx: *mut ndb_ctx, op: c_int, scalars: *const c_void, n_inputs: c_int, inputs: *const ndb_view, result: *mut c_int) -> c_int;
    pub fn ndb_zip_any(ctx: *mut ndb_ctx, op: c_int, scalars: *const c_void, n_inputs: c_int, inputs: *const ndb_view, result: *mut c_int) -> c_int;
    pub fn ndb_zip_fold_sum(ctx: *mut ndb_ctx, op: c_int, scalars: *const c_void, n_inputs: c_int, inputs: *const ndb_view, out_host: *mut c_void) -> c_int;
    pub fn ndb_zip_host(ctx: *mut ndb_ctx, op: c_int, scalars: *const c_void, n_parts: c_int, parts: *const ndb_view) -> c_int;
    pub fn ndb_reduce_all_dev(ctx: *mut ndb_ctx, redop: c_int, input: *const ndb_view, out_dev: *mut c_void) -> c_int;
    pub fn ndb_welford_axis(ctx: *mut ndb_ctx, axis: c_int, input: *const ndb_view, mean_out: *const ndb_view, m2_out: *const ndb_view) -> c_int;
    pub fn ndb_welford_merge(ctx: *mut ndb_ctx, n_shards: c_int, counts: *const i64, means: *const ndb_view, m2s: *const ndb_view, ddof: c_double, want_std: c_int, out: *const ndb_view) -> c_int;
    pub fn ndb_var(ctx: *mut ndb_ctx, input: *const ndb_view, ddof: c_double, want_std: c_int, out_host: *mut c_void) -> c_int;
    pub fn ndb_peer_window_create(ctx: *mut ndb_ctx, rank: c_int, world: c_int, slot_bytes: size_t, handle_out: *mut c_void) -> c_int;
    pub fn ndb_peer_window_attach(ctx: *mut ndb_ctx, handles: *const c_void) -> c_int;
    pub fn ndb_peer_window_destroy(ctx: *mut ndb_ctx) -> c_int;
    pub fn ndb_peer_allreduce(ctx: *mut ndb_ctx, redop: c_int, dtype: c_int, local_dev: *const c_void, out_dev: *mut c_void, n: i64) -> c_int;
    pub fn ndb_peer_allgather(ctx: *mut ndb_ctx, dtype: c_int, local_dev: *const c_void, out_dev: *mut c_void, n: i64) -> c_int;
    pub fn ndb_reduce_axis_peer(ctx: *mut ndb_ctx, redop: c_int, axis: c_int, input: *const ndb_view, out: *const ndb_view, ddof: c_double, n_global: i64) -> c_int;
    pub fn ndb_reduce_axis_peer_ok(ctx: *mut ndb_ctx, redop: c_int, axis: c_int, input: *const ndb_view, out: *const ndb_view, ddof: c_double, n_global: i64) -> c_int;
    pub fn ndb_reduce_all_peer(ctx: *mut ndb_ctx, redop: c_int, input: *const ndb_view, out_dev: *mut c_void) -> c_int;
    pub fn ndb_accumulate_axis(ctx: *mut ndb_ctx, op: c_int, axis: c_int, v: *const ndb_view) -> c_int;
    pub fn ndb_zip_fold(ctx: *mut ndb_ctx, op: c_int, scalars: *const c_void, n_inputs: c_int, inputs: *const ndb_view, redop: c_int, init: *const c_void, out_host: *mut c_void) -> c_int;
    pub fn ndb_zip_fold_write(ctx: *mut ndb_ctx, op: c_int, scalars: *const c_void, n_parts: c_int, parts: *const ndb_view, redop: c_int, out_dev: *mut c_void, across_ranks: c_int) -> c_int;
    pub fn ndb_zip_fold_while(ctx: *mut ndb_ctx, op: c_int, scalars: *const c_void, n_inputs: c_int, inputs: *const ndb_view, redop: c_int, init: *const c_void, stop_op: c_int, stop_value: *const c_void, out_host: *mut c_void, done: *mut c_int) -> c_int;
    pub fn ndb_zip_host_fold(ctx: *mut ndb_ctx, op: c_int, scalars: *const c_void, n_parts: c_int, parts: *const ndb_view, redop: c_int, out_host: *mut c_void) -> c_int;
    pub fn ndb_fill_counter_rng(ctx: *mut ndb_ctx, out: *const ndb_view, seed: u64, first_index: u64) -> c_int;
    pub fn ndb_graph_begin(ctx: *mut ndb_ctx) -> c_int;
    pub fn ndb_graph_end(ctx: *mut ndb_ctx, graph_out: *mut *mut c_void) -> c_int;
    pub fn ndb_graph_launch(ctx: *mut ndb_ctx, graph: *mut c_void) -> c_int;
    pub fn ndb_graph_destroy(graph: *mut c_void) -> c_int;
    pub fn ndb_comm_unique_id(id_out: *mut c_void) -> c_int;
    pub fn ndb_comm_init(ctx: *mut ndb_ctx, rank: c_int, world: c_int, id: *const c_void) -> c_int;
    pub fn ndb_comm_destroy(ctx: *mut ndb_ctx) -> c_int;
    pub fn ndb_broadcast(ctx: *mut ndb_ctx, v: *const ndb_view, src: c_int) -> c_int;
    pub fn ndb_gemm_rowsharded(ctx: *mut ndb_ctx, alpha: *const c_void, a_local: *const ndb_view, b: *const ndb_view, beta: *const c_void, c_local: *const ndb_view, src: c_int) -> c_int;
    pub fn ndb_reduce_axis_sharded(ctx: *mut ndb_ctx, redop: c_int, axis: c_int, input: *const ndb_view, out: *const ndb_view, ddof: c_double, n_global: i64) -> c_int;
    pub fn ndb_reduce_all_sharded(ctx: *mut ndb_ctx, redop: c_int, input: *const ndb_view, out_host: *mut c_void) -> c_int;
    // ndarray's own CBLAS boundary (src/linalg/impl_linalg.rs:466-491, :116-122); normally bound through cblas-sys
    pub fn cblas_sgemm(layout: c_int, transa: c_int, transb: c_int, m: c_int, n: c_int, k: c_int, alpha: f32, a: *const f32, lda: c_int, b: *const f32, ldb: c_int, beta: f32, c: *mut f32, ldc: c_int);
    pub fn cblas_dgemm(layout: c_int, transa: c_int, transb: c_int, m: c_int, n: c_int, k: c_int, alpha: f64, a: *const f64, lda: c_int, b: *const f64, ldb: c_int, beta: f64, c: *mut f64, ldc: c_int);
    pub fn cblas_sgemv(layout: c_int, trans: c_int, m: c_int, n: c_int, alpha: f32, a: *const f32, lda: c_int, x: *const f32, incx: c_int, beta: f32, y: *mut f32, incy: c_int);
    pub fn cblas_dgemv(layout: c_int, trans: c_int, m: c_int, n: c_int, alpha: f64, a: *const f64, lda: c_int, x: *const f64, incx: c_int, beta: f64, y: *mut f64, incy: c_int);
    pub fn cblas_sdot(n: c_int, x: *const f32, incx: c_int, y: *const f32, incy: c_int) -> f32;
    pub fn cblas_ddot(n: c_int, x: *const f64, incx: c_int, y: *const f64, incy: c_int) -> f64;
}
