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
}
