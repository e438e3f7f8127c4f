//! `ndarray-b200`: device arrays that keep ndarray's API surface.
//!
//! Shapes, strides, `Axis`, `Slice`/`s![]` and `Dimension` are ndarray's own types (re-exported), so a call
//! site changes by type alone:
//!
//! ```ignore
//! use ndarray_b200::{DeviceArray, Zip, Op, general_mat_mul};
//! use ndarray::{s, Axis};
//! let mut a = DeviceArray::<f32, _>::from_host(&host_a);      // was: host_a.clone()
//! let (b, c) = (DeviceArray::from_host(&host_b), DeviceArray::from_host(&host_c));
//! Zip::from(&mut a).and(&b).and(&c).for_each(Op::MulAddAssign); // was: .for_each(|a,&b,&c| *a = b*c + *a)
//! let s0 = a.sum_axis(Axis(0));
//! let v = b.slice(s![..;-2, .., ..]);                          // zero-copy view, identical (ptr, dim, strides)
//! general_mat_mul(1.0, &a2, &b2, 0.0, &mut c2);
//! ```
//!
//! Every method is one call into libndb200 (`ffi`).  There is no CPU fallback.
pub mod ffi;

use ndarray::{Axis, Dimension, IntoDimension, Ix1, Ix2, RemoveAxis, ShapeBuilder, Slice, SliceInfoElem};
use std::{ffi::CStr, marker::PhantomData, os::raw::c_void, sync::Arc};

/// Element types the library carries kernels for.
pub trait Elem: Copy + 'static { const DTYPE: i32; }
impl Elem for f32 { const DTYPE: i32 = 0; }
impl Elem for f64 { const DTYPE: i32 = 1; }
impl Elem for i32 { const DTYPE: i32 = 2; }
impl Elem for i64 { const DTYPE: i32 = 3; }

pub struct Context(pub(crate) *mut ffi::ndb_ctx);
unsafe impl Send for Context {}
unsafe impl Sync for Context {}
impl Context {
    pub fn new(device: i32) -> Arc<Context> {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::ndb_ctx_create(device, &mut h) });
        Arc::new(Context(h))
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { ffi::ndb_ctx_destroy(self.0); } } }

fn last_error() -> String { unsafe { CStr::from_ptr(ffi::ndb_last_error()) }.to_string_lossy().into_owned() }
/// ndarray panics on shape mismatch (src/zip/mod.rs:252, src/linalg/impl_linalg.rs:351); so do we, same text.
fn check(status: i32) { if status != ffi::NDB_OK { panic!("{}", last_error()); } }

struct Buffer { ctx: Arc<Context>, ptr: *mut c_void }
impl Drop for Buffer { fn drop(&mut self) { unsafe { ffi::ndb_free(self.ctx.0, self.ptr); } } }

/// Owned device array or view: the (ptr, dim, strides) triple of ndarray's ArrayBase (src/lib.rs:1304).
pub struct DeviceArray<A: Elem, D: Dimension> {
    buf: Arc<Buffer>,
    ptr: *mut A,          // logical first element
    dim: D,
    strides: D,           // signed element strides stored as usize, like ndarray
    _a: PhantomData<A>,
}
pub type DeviceView<'a, A, D> = &'a DeviceArray<A, D>;

impl<A: Elem, D: Dimension> DeviceArray<A, D> {
    pub fn uninit<Sh: ShapeBuilder<Dim = D>>(ctx: &Arc<Context>, shape: Sh) -> Self {
        let shape = shape.into_shape_with_order();
        let dim = shape.raw_dim().clone();
        let strides = if shape.is_c() { dim.default_strides() } else { dim.fortran_strides() };
        let mut p = std::ptr::null_mut();
        check(unsafe { ffi::ndb_alloc(ctx.0, dim.size() * std::mem::size_of::<A>(), &mut p) });
        DeviceArray { buf: Arc::new(Buffer { ctx: ctx.clone(), ptr: p }), ptr: p as *mut A, dim, strides, _a: PhantomData }
    }
    pub fn from_host(ctx: &Arc<Context>, a: &ndarray::Array<A, D>) -> Self {
        let std_a = a.as_standard_layout();
        let out = Self::uninit(ctx, a.raw_dim());
        check(unsafe { ffi::ndb_h2d(ctx.0, out.ptr as *mut c_void, std_a.as_ptr() as *const c_void, a.len() * std::mem::size_of::<A>()) });
        check(unsafe { ffi::ndb_sync(ctx.0) });
        out
    }
    pub fn raw_dim(&self) -> D { self.dim.clone() }
    pub fn strides(&self) -> &[isize] { unsafe { std::slice::from_raw_parts(self.strides.slice().as_ptr() as *const isize, self.dim.ndim()) } }
    pub(crate) fn view(&self) -> ffi::ndb_view {
        let mut v = ffi::ndb_view { ptr: self.ptr as *mut c_void, ndim: self.dim.ndim() as i32, dtype: A::DTYPE,
                                    dim: [0; ffi::NDB_MAX_NDIM], stride: [0; ffi::NDB_MAX_NDIM] };
        for (i, (&d, &s)) in self.dim.slice().iter().zip(self.strides.slice()).enumerate() { v.dim[i] = d as i64; v.stride[i] = s as isize as i64; }
        v
    }
    fn with_view(&self, v: ffi::ndb_view) -> Self {
        let mut dim = self.dim.clone(); let mut st = self.strides.clone();
        for i in 0..dim.ndim() { dim[i] = v.dim[i] as usize; st[i] = v.stride[i] as isize as usize; }
        DeviceArray { buf: self.buf.clone(), ptr: v.ptr as *mut A, dim, strides: st, _a: PhantomData }
    }
    /// `slice_axis` (src/impl_methods.rs:689): do_slice runs in the library, bit-identical to ndarray's.
    pub fn slice_axis(&self, axis: Axis, s: Slice) -> Self {
        let mut v = self.view();
        check(unsafe { ffi::ndb_view_slice_axis(&mut v, axis.index() as i32, s.start as i64, s.end.unwrap_or(0) as i64, s.end.is_some() as i32, s.step as i64) });
        self.with_view(v)
    }
    /// `slice(s![..])` for the same-rank case (every element a `Slice`).
    pub fn slice(&self, info: &[SliceInfoElem]) -> Self {
        let mut out = self.with_view(self.view());
        for (ax, e) in info.iter().enumerate() {
            if let SliceInfoElem::Slice { start, end, step } = *e { out = out.slice_axis(Axis(ax), Slice { start, end, step }); }
            else { unimplemented!("index / NewAxis elements change the rank: use index_axis / insert_axis"); }
        }
        out
    }
    pub fn reversed_axes(&self) -> Self {
        let n = self.dim.ndim(); let perm: Vec<i32> = (0..n as i32).rev().collect();
        let mut out = self.view();
        check(unsafe { ffi::ndb_view_permute(&self.view(), perm.as_ptr(), &mut out) });
        self.with_view(out)
    }
    pub fn t(&self) -> Self { self.reversed_axes() }
    pub fn invert_axis(&self, axis: Axis) -> Self { let mut v = self.view(); check(unsafe { ffi::ndb_view_invert_axis(&mut v, axis.index() as i32) }); self.with_view(v) }
    /// `broadcast` (src/impl_methods.rs:2371): None when the shapes are incompatible.
    pub fn broadcast<E: IntoDimension>(&self, dim: E) -> Option<DeviceArray<A, E::Dim>> {
        let dim = dim.into_dimension(); let d: Vec<i64> = dim.slice().iter().map(|&x| x as i64).collect();
        let mut out = self.view();
        let st = unsafe { ffi::ndb_view_broadcast(&self.view(), d.len() as i32, d.as_ptr(), &mut out) };
        if st == ffi::NDB_ERR_INCOMPATIBLE_SHAPE { return None; }
        check(st);
        let mut strides = dim.clone();
        for i in 0..dim.ndim() { strides[i] = out.stride[i] as isize as usize; }
        Some(DeviceArray { buf: self.buf.clone(), ptr: out.ptr as *mut A, dim, strides, _a: PhantomData })
    }
    pub fn sum(&self) -> A where A: Default {
        let mut out = A::default();
        check(unsafe { ffi::ndb_reduce_all(self.buf.ctx.0, 0, &self.view(), &mut out as *mut A as *mut c_void) });
        out
    }
}

impl<A: Elem, D: Dimension + RemoveAxis> DeviceArray<A, D> {
    fn reduce_axis(&self, redop: i32, axis: Axis, ddof: f64) -> Option<DeviceArray<A, D::Smaller>> {
        let out = DeviceArray::<A, D::Smaller>::uninit(&self.buf.ctx, self.dim.remove_axis(axis));
        let st = unsafe { ffi::ndb_reduce_axis(self.buf.ctx.0, redop, axis.index() as i32, &self.view(), &out.view(), ddof) };
        if st == ffi::NDB_ERR_EMPTY { return None; }
        check(st);
        Some(out)
    }
    pub fn sum_axis(&self, axis: Axis) -> DeviceArray<A, D::Smaller> { self.reduce_axis(0, axis, 0.0).unwrap() }
    pub fn product_axis(&self, axis: Axis) -> DeviceArray<A, D::Smaller> { self.reduce_axis(1, axis, 0.0).unwrap() }
    pub fn mean_axis(&self, axis: Axis) -> Option<DeviceArray<A, D::Smaller>> { self.reduce_axis(4, axis, 0.0) }
    pub fn var_axis(&self, axis: Axis, ddof: f64) -> DeviceArray<A, D::Smaller> { self.reduce_axis(5, axis, ddof).unwrap() }
    pub fn std_axis(&self, axis: Axis, ddof: f64) -> DeviceArray<A, D::Smaller> { self.reduce_axis(6, axis, ddof).unwrap() }
}

/// The closures ndarray passes to Zip, as an op vocabulary (closures cannot cross the FFI).
#[derive(Clone, Copy)]
pub enum Op { Assign, Add, Sub, Mul, Div, AddAssign, SubAssign, MulAssign, DivAssign, MulAddAssign }

/// `Zip::from(a).and(b).and(c)` (src/zip/mod.rs:193): producers of identical shape, first one written.
pub struct Zip<A: Elem> { ctx: *mut ffi::ndb_ctx, parts: Vec<ffi::ndb_view>, _a: PhantomData<A> }
impl<A: Elem> Zip<A> {
    pub fn from<D: Dimension>(p: &mut DeviceArray<A, D>) -> Self { Zip { ctx: p.buf.ctx.0, parts: vec![p.view()], _a: PhantomData } }
    pub fn and<D: Dimension>(mut self, p: &DeviceArray<A, D>) -> Self { self.parts.push(p.view()); self }
    pub fn for_each(self, op: Op) {
        let p = &self.parts;
        let (code, parts): (i32, Vec<ffi::ndb_view>) = match op {
            Op::Assign => (0, vec![p[0], p[1]]),
            Op::Add => (2, vec![p[0], p[1], p[2]]), Op::Sub => (3, vec![p[0], p[1], p[2]]),
            Op::Mul => (4, vec![p[0], p[1], p[2]]), Op::Div => (5, vec![p[0], p[1], p[2]]),
            Op::AddAssign => (2, vec![p[0], p[0], p[1]]), Op::SubAssign => (3, vec![p[0], p[0], p[1]]),
            Op::MulAssign => (4, vec![p[0], p[0], p[1]]), Op::DivAssign => (5, vec![p[0], p[0], p[1]]),
            Op::MulAddAssign => (7, vec![p[0], p[1], p[2], p[0]]),
        };
        check(unsafe { ffi::ndb_zip(self.ctx, code, std::ptr::null(), parts.len() as i32, parts.as_ptr()) });
    }
    pub fn par_for_each(self, op: Op) { self.for_each(op) }
}

/// `general_mat_mul(alpha, &a, &b, beta, &mut c)` (src/linalg/impl_linalg.rs:633).
pub fn general_mat_mul<A: Elem>(alpha: A, a: &DeviceArray<A, Ix2>, b: &DeviceArray<A, Ix2>, beta: A, c: &mut DeviceArray<A, Ix2>) {
    check(unsafe { ffi::ndb_gemm(c.buf.ctx.0, &alpha as *const A as *const c_void, &a.view(), &b.view(), &beta as *const A as *const c_void, &c.view()) });
}
/// 1-d dot (src/linalg/impl_linalg.rs:74).
pub fn dot<A: Elem + Default>(x: &DeviceArray<A, Ix1>, y: &DeviceArray<A, Ix1>) -> A {
    let mut out = A::default();
    check(unsafe { ffi::ndb_dot(x.buf.ctx.0, &x.view(), &y.view(), &mut out as *mut A as *mut c_void) });
    out
}

// ---- the rest of SURVEY §8f: predicates / fold over a Zip, casts, scans, matrix-vector ----------------------------

/// Comparison predicates for `Zip::all` / `Zip::any` (src/zip/mod.rs:674-696); closures cannot cross the FFI.
pub enum Pred { Eq = 25, Ne = 26, Lt = 27, Le = 28, Gt = 29, Ge = 30 }

impl<A: Elem> Zip<A> {
    /// `Zip::from(&a).and(&b).all(|&a, &b| a < b)`: the first two producers are compared; vacuously true when empty.
    pub fn all(self, pred: Pred) -> bool {
        let mut r = 0i32;
        check(unsafe { ffi::ndb_zip_all(self.ctx, pred as i32, std::ptr::null(), self.parts.len() as i32, self.parts.as_ptr(), &mut r) });
        r != 0
    }
    pub fn any(self, pred: Pred) -> bool {
        let mut r = 0i32;
        check(unsafe { ffi::ndb_zip_any(self.ctx, pred as i32, std::ptr::null(), self.parts.len() as i32, self.parts.as_ptr(), &mut r) });
        r != 0
    }
    /// `Zip::fold(0, |acc, &a, &b| acc + a * b)` for the additive folds (op = Mul / Add / Sub of the two producers).
    pub fn fold_sum(self, op: Op) -> A where A: Default {
        let code = match op { Op::Mul => 4, Op::Add => 2, Op::Sub => 3, _ => panic!("fold_sum: op must be Mul, Add or Sub") };
        let mut out = A::default();
        check(unsafe { ffi::ndb_zip_fold_sum(self.ctx, code, std::ptr::null(), self.parts.len() as i32, self.parts.as_ptr(), &mut out as *mut A as *mut c_void) });
        out
    }
}

impl<A: Elem, D: Dimension> DeviceArray<A, D> {
    /// `a.mapv(|x| x as B)` (Rust `as` semantics: saturating float -> int, NaN -> 0), src/impl_methods.rs:2934.
    pub fn mapv_cast<B: Elem>(&self) -> DeviceArray<B, D> {
        let out = DeviceArray::<B, D>::uninit(&self.buf.ctx, self.dim.clone());
        check(unsafe { ffi::ndb_cast(self.buf.ctx.0, &out.view(), &self.view()) });
        out
    }
    /// `accumulate_axis_inplace(axis, |&prev, cur| *cur = *cur + prev)` / `... * prev` (src/impl_methods.rs:3243): one scan launch.
    pub fn accumulate_axis_inplace(&mut self, axis: Axis, op: Op) {
        let code = match op { Op::Add | Op::AddAssign => 2, Op::Mul | Op::MulAssign => 4, _ => panic!("accumulate: op must be Add or Mul") };
        check(unsafe { ffi::ndb_accumulate_axis(self.buf.ctx.0, code, axis.index() as i32, &self.view()) });
    }
    /// whole-array `var(ddof)` / `std(ddof)` (src/numeric/impl_numeric.rs:180-249)
    pub fn var(&self, ddof: f64) -> A where A: Default {
        let mut out = A::default();
        check(unsafe { ffi::ndb_var(self.buf.ctx.0, &self.view(), ddof, 0, &mut out as *mut A as *mut c_void) });
        out
    }
}

/// `general_mat_vec_mul(alpha, &a, &x, beta, &mut y)` (src/linalg/impl_linalg.rs:657).
pub fn general_mat_vec_mul<A: Elem>(alpha: A, a: &DeviceArray<A, Ix2>, x: &DeviceArray<A, Ix1>, beta: A, y: &mut DeviceArray<A, Ix1>) {
    check(unsafe { ffi::ndb_gemv(y.buf.ctx.0, &alpha as *const A as *const c_void, &a.view(), &x.view(), &beta as *const A as *const c_void, &y.view()) });
}

// ---- multi-GPU: one Context per GPU (one process per GPU), arrays sharded along Axis(0) ------------------------------

/// Balanced, contiguous, rank-ordered partition of an axis of length `n` (what `Zip::split` does recursively on one host).
pub fn shard_bounds(n: usize, rank: usize, world: usize) -> (usize, usize) { (n * rank / world, n * (rank + 1) / world) }

/// Peer windows over NVLink (include/ndb200.h "multi-GPU exchange"): created once per context; `exchange_handles` is any
/// host-side all-gather of 64-byte blobs the application already has (MPI, a socket, a shared file).
pub struct PeerGroup { ctx: Arc<Context>, pub rank: usize, pub world: usize }
impl PeerGroup {
    pub fn new(ctx: &Arc<Context>, rank: usize, world: usize, slot_bytes: usize, exchange_handles: impl FnOnce([u8; 64]) -> Vec<[u8; 64]>) -> Self {
        let mut h = [0u8; 64];
        check(unsafe { ffi::ndb_peer_window_create(ctx.0, rank as i32, world as i32, slot_bytes, h.as_mut_ptr() as *mut c_void) });
        let all = exchange_handles(h);
        assert_eq!(all.len(), world);
        check(unsafe { ffi::ndb_peer_window_attach(ctx.0, all.as_ptr() as *const c_void) });
        PeerGroup { ctx: ctx.clone(), rank, world }
    }
    /// `sum()` over all shards: the kernel that finishes this rank's reduction also exchanges the scalar (no second launch).
    pub fn global_sum<A: Elem + Default, D: Dimension>(&self, shard: &DeviceArray<A, D>) -> A {
        let out = DeviceArray::<A, ndarray::Ix0>::uninit(&self.ctx, ());
        check(unsafe { ffi::ndb_reduce_all_peer(self.ctx.0, 0, &shard.view(), out.ptr as *mut c_void) });
        let mut host = A::default();
        check(unsafe { ffi::ndb_d2h(self.ctx.0, &mut host as *mut A as *mut c_void, out.ptr as *const c_void, std::mem::size_of::<A>()) });
        check(unsafe { ffi::ndb_sync(self.ctx.0) });
        host
    }
    /// `sum_axis` / `mean_axis` / `var_axis` / `std_axis(Axis(0))` of a row-sharded array: one fused launch per rank when
    /// the layout allows (all ranks must have agreed on that once, `ndb_reduce_axis_peer_ok`), else reduce + exchange.
    pub fn reduce_axis0<A: Elem, D: Dimension + RemoveAxis>(&self, redop: i32, shard: &DeviceArray<A, D>, ddof: f64, n_global: usize, fused: bool) -> DeviceArray<A, D::Smaller> {
        let out = DeviceArray::<A, D::Smaller>::uninit(&self.ctx, shard.dim.remove_axis(Axis(0)));
        if fused {
            check(unsafe { ffi::ndb_reduce_axis_peer(self.ctx.0, redop, 0, &shard.view(), &out.view(), ddof, n_global as i64) });
        } else {
            assert!(redop <= 3, "mean/var over shards need the fused path or ndb_welford_axis + ndb_peer_allgather + ndb_welford_merge");
            check(unsafe { ffi::ndb_reduce_axis(self.ctx.0, redop, 0, &shard.view(), &out.view(), 0.0) });
            let n: usize = out.dim.size();
            check(unsafe { ffi::ndb_peer_allreduce(self.ctx.0, redop, A::DTYPE, out.ptr as *const c_void, out.ptr as *mut c_void, n as i64) });
        }
        out
    }
}
impl Drop for PeerGroup { fn drop(&mut self) { unsafe { ffi::ndb_peer_window_destroy(self.ctx.0); } } }
