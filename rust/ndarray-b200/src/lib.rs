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

// =====================================================================================================================
// Round 2: the rest of the surface a call site meets (8e through the C ABI).
// Still uncompiled here (no rustc in the image); every method is one call of an entry point the GPU tests exercise
// through ctypes, and tests/test_host_logic.py checks that src/ffi.rs declares exactly the header's symbols.

/// Associative accumulators for `Zip::fold` / `fold_axis` (closures cannot cross the FFI; src/zip/mod.rs:622,
/// src/impl_methods.rs:3130): the numbering is `ndb_redop`.
#[derive(Clone, Copy)]
pub enum Fold { Sum = 0, Prod = 1, Min = 2, Max = 3 }

/// Element maps a `Zip` can apply (`ndb_op`): what `azip!` bodies reduce to.  `Op` above keeps the assign forms.
#[derive(Clone, Copy)]
pub enum Map { Copy = 0, Add = 2, Sub = 3, Mul = 4, Div = 5, Rem = 6, MulAdd = 7, Min = 20, Max = 21 }

impl<A: Elem, D: Dimension> DeviceArray<A, D> {
    pub fn ctx(&self) -> &Arc<Context> { &self.buf.ctx }
    pub fn shape(&self) -> &[usize] { self.dim.slice() }
    pub fn len(&self) -> usize { self.dim.size() }
    pub fn is_standard_layout(&self) -> bool { unsafe { ffi::ndb_view_layout(&self.view()) & 1 != 0 } }
    pub fn is_contiguous(&self) -> bool { unsafe { ffi::ndb_view_is_contiguous(&self.view()) != 0 } }

    /// `permuted_axes` (src/impl_methods.rs:2522): out axis i = in axis axes[i].
    pub fn permuted_axes<T: IntoDimension<Dim = D>>(&self, axes: T) -> Self {
        let axes = axes.into_dimension();
        let perm: Vec<i32> = axes.slice().iter().map(|&a| a as i32).collect();
        let mut out = self.view();
        check(unsafe { ffi::ndb_view_permute(&self.view(), perm.as_ptr(), &mut out) });
        self.with_view(out)
    }
    /// `swap_axes` (src/impl_methods.rs:2487), in place like ndarray's.
    pub fn swap_axes(&mut self, ax: usize, bx: usize) {
        self.dim.slice_mut().swap(ax, bx);
        self.strides.slice_mut().swap(ax, bx);
    }
    /// `split_at` (src/impl_raw_views.rs:113-135): the sharding primitive.
    pub fn split_at(&self, axis: Axis, index: usize) -> (Self, Self) {
        let (mut l, mut r) = (self.view(), self.view());
        check(unsafe { ffi::ndb_view_split_at(&self.view(), axis.index() as i32, index as i64, &mut l, &mut r) });
        (self.with_view(l), self.with_view(r))
    }
    /// This rank's slab along Axis(0) (two `split_at`s; stride-0 operands are replicated as their stride says).
    pub fn shard_axis0(&self, rank: usize, world: usize) -> Self {
        let (lo, hi) = shard_bounds(self.dim[0], rank, world);
        let (_, right) = self.split_at(Axis(0), lo);
        right.split_at(Axis(0), hi - lo).0
    }
    /// `to_owned` in standard layout (one Zip copy on the device).
    pub fn to_owned(&self) -> Self {
        let out = Self::uninit(&self.buf.ctx, self.dim.clone());
        let parts = [out.view(), self.view()];
        check(unsafe { ffi::ndb_zip(self.buf.ctx.0, Map::Copy as i32, std::ptr::null(), 2, parts.as_ptr()) });
        out
    }
    /// Copy back to an `ndarray::Array` (standard layout).
    pub fn to_host(&self) -> ndarray::Array<A, D> where A: Default {
        let dense = if self.is_standard_layout() { self.with_view(self.view()) } else { self.to_owned() };
        let mut host = vec![A::default(); self.len()];
        check(unsafe { ffi::ndb_d2h(self.buf.ctx.0, host.as_mut_ptr() as *mut c_void, dense.ptr as *const c_void, host.len() * std::mem::size_of::<A>()) });
        check(unsafe { ffi::ndb_sync(self.buf.ctx.0) });
        ndarray::Array::from_shape_vec(self.dim.clone(), host).unwrap()
    }
    /// `fill` (src/impl_methods.rs:2819).
    pub fn fill(&mut self, x: A) {
        let parts = [self.view()];
        check(unsafe { ffi::ndb_zip(self.buf.ctx.0, 1, &x as *const A as *const c_void, 1, parts.as_ptr()) });
    }
    /// `assign` with broadcasting of the right-hand side (src/impl_methods.rs:2796).
    pub fn assign<E: Dimension>(&mut self, rhs: &DeviceArray<A, E>) {
        let r = rhs.broadcast(self.dim.clone()).unwrap_or_else(|| panic!("ndarray: could not broadcast array from shape: {:?} to: {:?}", rhs.shape(), self.shape()));
        let parts = [self.view(), r.view()];
        check(unsafe { ffi::ndb_zip(self.buf.ctx.0, Map::Copy as i32, std::ptr::null(), 2, parts.as_ptr()) });
    }
    /// `zip_mut_with(&rhs, |a, &b| *a = a <op> b)` with broadcasting (src/impl_methods.rs:2825-2896).
    pub fn zip_mut_with<E: Dimension>(&mut self, rhs: &DeviceArray<A, E>, op: Map) {
        let r = rhs.broadcast(self.dim.clone()).unwrap_or_else(|| panic!("ndarray: could not broadcast array from shape: {:?} to: {:?}", rhs.shape(), self.shape()));
        let parts = [self.view(), self.view(), r.view()];
        check(unsafe { ffi::ndb_zip(self.buf.ctx.0, op as i32, std::ptr::null(), 3, parts.as_ptr()) });
    }
    /// `scaled_add(alpha, &rhs)` (src/linalg/impl_linalg.rs:399): self += alpha * rhs.
    pub fn scaled_add<E: Dimension>(&mut self, alpha: A, rhs: &DeviceArray<A, E>) {
        let r = rhs.broadcast(self.dim.clone()).expect("scaled_add: shapes do not broadcast");
        let parts = [self.view(), self.view(), r.view()];
        check(unsafe { ffi::ndb_zip(self.buf.ctx.0, 16, &alpha as *const A as *const c_void, 3, parts.as_ptr()) });
    }
    fn map_scalar(&self, code: i32, s: A) -> Self {
        let out = Self::uninit(&self.buf.ctx, self.dim.clone());
        let parts = [out.view(), self.view()];
        check(unsafe { ffi::ndb_zip(self.buf.ctx.0, code, &s as *const A as *const c_void, 2, parts.as_ptr()) });
        out
    }
    fn map_binary<E: Dimension>(&self, rhs: &DeviceArray<A, E>, code: i32) -> Self {
        // same-shape fast path; otherwise the right operand is broadcast to ours (co-broadcast of both sides is
        // `broadcast()` on each first, src/impl_ops.rs:123-151)
        let r = rhs.broadcast(self.dim.clone()).unwrap_or_else(|| panic!("ndarray: could not broadcast array from shape: {:?} to: {:?}", rhs.shape(), self.shape()));
        let out = Self::uninit(&self.buf.ctx, self.dim.clone());
        let parts = [out.view(), self.view(), r.view()];
        check(unsafe { ffi::ndb_zip(self.buf.ctx.0, code, std::ptr::null(), 3, parts.as_ptr()) });
        out
    }
    pub fn product(&self) -> A where A: Default {
        let mut out = A::default();
        check(unsafe { ffi::ndb_reduce_all(self.buf.ctx.0, 1, &self.view(), &mut out as *mut A as *mut c_void) });
        out
    }
    pub fn std(&self, ddof: f64) -> A where A: Default {
        let mut out = A::default();
        check(unsafe { ffi::ndb_var(self.buf.ctx.0, &self.view(), ddof, 1, &mut out as *mut A as *mut c_void) });
        out
    }
}

impl<A: Elem, D: Dimension + RemoveAxis> DeviceArray<A, D> {
    /// `index_axis` (src/impl_methods.rs:1009): fix `axis` at `index`, the axis disappears.
    pub fn index_axis(&self, axis: Axis, index: usize) -> DeviceArray<A, D::Smaller> {
        let mut v = self.view();
        check(unsafe { ffi::ndb_view_index_axis(&mut v, axis.index() as i32, index as i64) });
        let mut dim = self.dim.remove_axis(axis);
        let mut strides = dim.clone();
        for i in 0..dim.ndim() { dim[i] = v.dim[i] as usize; strides[i] = v.stride[i] as isize as usize; }
        DeviceArray { buf: self.buf.clone(), ptr: v.ptr as *mut A, dim, strides, _a: PhantomData }
    }
    /// `fold_axis(axis, init, |&acc, &x| acc <fold> x)` for the associative folds (src/impl_methods.rs:3130-3141):
    /// the tree reduction, then `init` is combined once (identical for an associative, commutative closure).
    pub fn fold_axis(&self, axis: Axis, init: A, fold: Fold) -> DeviceArray<A, D::Smaller> {
        let mut out = self.reduce_axis(fold as i32, axis, 0.0).unwrap_or_else(|| {
            let mut o = DeviceArray::<A, D::Smaller>::uninit(&self.buf.ctx, self.dim.remove_axis(axis));
            o.fill(init);
            o
        });
        if self.dim[axis.index()] == 0 { return out; }
        match fold {
            Fold::Sum => { let parts = [out.view(), out.view()]; check(unsafe { ffi::ndb_zip(self.buf.ctx.0, 8, &init as *const A as *const c_void, 2, parts.as_ptr()) }); }
            Fold::Prod => { let parts = [out.view(), out.view()]; check(unsafe { ffi::ndb_zip(self.buf.ctx.0, 10, &init as *const A as *const c_void, 2, parts.as_ptr()) }); }
            Fold::Min | Fold::Max => {
                let mut ini = DeviceArray::<A, D::Smaller>::uninit(&self.buf.ctx, out.dim.clone());
                ini.fill(init);
                let parts = [out.view(), ini.view(), out.view()];
                let code = if matches!(fold, Fold::Min) { Map::Min } else { Map::Max } as i32;
                check(unsafe { ffi::ndb_zip(self.buf.ctx.0, code, std::ptr::null(), 3, parts.as_ptr()) });
            }
        }
        out
    }
}

impl<A: Elem, D: Dimension> DeviceArray<A, D> {
    /// `insert_axis` (src/impl_methods.rs:2700): a new axis of length 1 (stride 1, like ndarray).
    pub fn insert_axis(&self, axis: Axis) -> DeviceArray<A, D::Larger> {
        let dim = self.dim.insert_axis(axis);
        let mut strides = dim.clone();
        let (mut j, ax) = (0, axis.index());
        for i in 0..dim.ndim() {
            if i == ax { strides[i] = 1; } else { strides[i] = self.strides[j]; j += 1; }
        }
        DeviceArray { buf: self.buf.clone(), ptr: self.ptr, dim, strides, _a: PhantomData }
    }
}

// ---- operators (src/impl_ops.rs:53-338, 348-405, 597-712): `&a + &b`, `&a * 2.0`, `a += &b`, `-&a` -----------------
macro_rules! impl_binary_op {
    ($trt:ident, $mth:ident, $code:expr, $code_s:expr, $asg:ident, $asg_mth:ident) => {
        impl<'a, A: Elem, D: Dimension, E: Dimension> std::ops::$trt<&'a DeviceArray<A, E>> for &'a DeviceArray<A, D> {
            type Output = DeviceArray<A, D>;
            fn $mth(self, rhs: &'a DeviceArray<A, E>) -> DeviceArray<A, D> { self.map_binary(rhs, $code) }
        }
        impl<'a, A: Elem, D: Dimension> std::ops::$trt<A> for &'a DeviceArray<A, D> {
            type Output = DeviceArray<A, D>;
            fn $mth(self, rhs: A) -> DeviceArray<A, D> { self.map_scalar($code_s, rhs) }
        }
        impl<'a, A: Elem, D: Dimension, E: Dimension> std::ops::$asg<&'a DeviceArray<A, E>> for DeviceArray<A, D> {
            fn $asg_mth(&mut self, rhs: &'a DeviceArray<A, E>) {
                let r = rhs.broadcast(self.dim.clone()).unwrap_or_else(|| panic!("ndarray: could not broadcast array from shape: {:?} to: {:?}", rhs.shape(), self.shape()));
                let parts = [self.view(), self.view(), r.view()];
                check(unsafe { ffi::ndb_zip(self.buf.ctx.0, $code, std::ptr::null(), 3, parts.as_ptr()) });
            }
        }
        impl<A: Elem, D: Dimension> std::ops::$asg<A> for DeviceArray<A, D> {
            fn $asg_mth(&mut self, rhs: A) {
                let parts = [self.view(), self.view()];
                check(unsafe { ffi::ndb_zip(self.buf.ctx.0, $code_s, &rhs as *const A as *const c_void, 2, parts.as_ptr()) });
            }
        }
    };
}
impl_binary_op!(Add, add, 2, 8, AddAssign, add_assign);
impl_binary_op!(Sub, sub, 3, 9, SubAssign, sub_assign);
impl_binary_op!(Mul, mul, 4, 10, MulAssign, mul_assign);
impl_binary_op!(Div, div, 5, 11, DivAssign, div_assign);
impl_binary_op!(Rem, rem, 6, 12, RemAssign, rem_assign);
impl<'a, A: Elem, D: Dimension> std::ops::Neg for &'a DeviceArray<A, D> {
    type Output = DeviceArray<A, D>;
    fn neg(self) -> DeviceArray<A, D> {
        let out = DeviceArray::<A, D>::uninit(&self.buf.ctx, self.dim.clone());
        let parts = [out.view(), self.view()];
        check(unsafe { ffi::ndb_zip(self.buf.ctx.0, 15, std::ptr::null(), 2, parts.as_ptr()) });
        out
    }
}

// ---- Zip: the remaining methods ------------------------------------------------------------------------------------
impl<A: Elem> Zip<A> {
    /// `Zip::from(&a)` for read-only traversals (`fold`, `all`, `map_collect`): no producer is written.
    pub fn from_ref<D: Dimension>(p: &DeviceArray<A, D>) -> Self { Zip { ctx: p.buf.ctx.0, parts: vec![p.view()], _a: PhantomData } }
    /// `and_broadcast` (src/zip/mod.rs:738-745): the producer is broadcast to the Zip's shape first.
    pub fn and_broadcast<D: Dimension>(mut self, p: &DeviceArray<A, D>) -> Self {
        let first = self.parts[0];
        let mut out = p.view();
        let st = unsafe { ffi::ndb_view_broadcast(&p.view(), first.ndim, first.dim.as_ptr(), &mut out) };
        if st != ffi::NDB_OK { panic!("ndarray: could not broadcast array from shape: {:?} to: {:?}", p.shape(), &first.dim[..first.ndim as usize]); }
        self.parts.push(out);
        self
    }
    /// `map_collect(|&a, &b| a <op> b)` (src/zip/mod.rs:764-789): fresh output, F-order iff the inputs prefer it.
    pub fn map_collect<D: Dimension>(self, shape_of: &DeviceArray<A, D>, op: Map) -> DeviceArray<A, D> {
        let f = unsafe { ffi::ndb_zip_prefer_f(self.parts.len() as i32, self.parts.as_ptr()) } != 0;
        let ctx = shape_of.buf.ctx.clone();
        let out = if f { DeviceArray::<A, D>::uninit(&ctx, shape_of.dim.clone().f()) } else { DeviceArray::<A, D>::uninit(&ctx, shape_of.dim.clone()) };
        let mut parts = vec![out.view()];
        parts.extend_from_slice(&self.parts);
        check(unsafe { ffi::ndb_zip(self.ctx, op as i32, std::ptr::null(), parts.len() as i32, parts.as_ptr()) });
        out
    }
    pub fn par_map_collect<D: Dimension>(self, shape_of: &DeviceArray<A, D>, op: Map) -> DeviceArray<A, D> { self.map_collect(shape_of, op) }
    /// `map_assign_into(&mut into, |..| ..)` (src/zip/mod.rs:796-804): panics with Zip's message on a shape mismatch.
    pub fn map_assign_into<D: Dimension>(self, into: &mut DeviceArray<A, D>, op: Map) {
        let mut parts = vec![into.view()];
        parts.extend_from_slice(&self.parts);
        check(unsafe { ffi::ndb_zip(self.ctx, op as i32, std::ptr::null(), parts.len() as i32, parts.as_ptr()) });
    }
    pub fn par_map_assign_into<D: Dimension>(self, into: &mut DeviceArray<A, D>, op: Map) { self.map_assign_into(into, op) }
    /// `fold(init, |acc, &a, &b| acc <fold> (a <op> b))` / `par_fold` for associative accumulators (src/zip/mod.rs:622).
    pub fn fold(self, init: A, op: Map, fold: Fold) -> A where A: Default {
        let mut out = A::default();
        check(unsafe { ffi::ndb_zip_fold(self.ctx, op as i32, std::ptr::null(), self.parts.len() as i32, self.parts.as_ptr(), fold as i32,
                                         &init as *const A as *const c_void, &mut out as *mut A as *mut c_void) });
        out
    }
    /// `Zip::from(&mut a).and(&b).and(&c).fold(acc0, |acc, a, &b, &c| { *a = b <op> c ..; acc <fold> *a })`: the first producer
    /// is written and the written values are folded in the same pass (`ndb_zip_fold_write`; one fused kernel when dense).
    pub fn for_each_fold(self, op: Op, fold: Fold) -> A where A: Default {
        let p = &self.parts;
        let (code, parts): (i32, Vec<ffi::ndb_view>) = match op {
            Op::Assign => (0, vec![p[0], p[1]]),
            Op::Add => (2, vec![p[0], p[1], p[2]]), Op::Sub => (3, vec![p[0], p[1], p[2]]),
            Op::Mul => (4, vec![p[0], p[1], p[2]]), Op::Div => (5, vec![p[0], p[1], p[2]]),
            Op::AddAssign => (2, vec![p[0], p[0], p[1]]), Op::SubAssign => (3, vec![p[0], p[0], p[1]]),
            Op::MulAssign => (4, vec![p[0], p[0], p[1]]), Op::DivAssign => (5, vec![p[0], p[0], p[1]]),
            Op::MulAddAssign => (7, vec![p[0], p[1], p[2], p[0]]),
        };
        let mut dev = std::ptr::null_mut();
        check(unsafe { ffi::ndb_alloc(self.ctx, 16, &mut dev) });
        check(unsafe { ffi::ndb_zip_fold_write(self.ctx, code, std::ptr::null(), parts.len() as i32, parts.as_ptr(), fold as i32, dev, 0) });
        let mut out = A::default();
        check(unsafe { ffi::ndb_d2h(self.ctx, &mut out as *mut A as *mut c_void, dev, std::mem::size_of::<A>()) });
        check(unsafe { ffi::ndb_sync(self.ctx) });
        check(unsafe { ffi::ndb_free(self.ctx, dev) });
        out
    }
    /// `fold_while(init, |acc, ..| if v <stop> s { FoldWhile::Done(acc) } else { FoldWhile::Continue(acc <fold> v) })`
    /// (src/zip/mod.rs:640-655) in Zip's traversal order; `FoldWhile` is ndarray's own type.
    pub fn fold_while(self, init: A, op: Map, fold: Fold, stop: Pred, stop_value: A) -> ndarray::FoldWhile<A> where A: Default {
        let (mut out, mut done) = (A::default(), 0i32);
        check(unsafe { ffi::ndb_zip_fold_while(self.ctx, op as i32, std::ptr::null(), self.parts.len() as i32, self.parts.as_ptr(), fold as i32,
                                               &init as *const A as *const c_void, stop as i32, &stop_value as *const A as *const c_void,
                                               &mut out as *mut A as *mut c_void, &mut done) });
        if done != 0 { ndarray::FoldWhile::Done(out) } else { ndarray::FoldWhile::Continue(out) }
    }
    /// The kernel class the planner picks (debugging / tests).
    pub fn plan(&self, op: Map) -> ffi::ndb_zip_plan_t {
        let mut pl = std::mem::MaybeUninit::<ffi::ndb_zip_plan_t>::zeroed();
        check(unsafe { ffi::ndb_zip_plan(op as i32, self.parts.len() as i32, self.parts.as_ptr(), pl.as_mut_ptr()) });
        unsafe { pl.assume_init() }
    }
}

/// `azip!`-shaped sugar (src/zip/zipmacro.rs:98, src/parallel/zipmacro.rs:53).  A closure body cannot cross the FFI, so the
/// body is one of the op vocabulary's forms; the producer list reads like ndarray's:
///
/// ```ignore
/// azip!((a in &mut a, b in &b, c in &c) a = b * c + a);   // MUL_ADD, a read and written
/// azip!((a in &mut a, b in &b, c in &c) a = b + c);
/// azip!((a in &mut a, b in &b) a += b);
/// par_azip!((a in &mut a, b in &b) a = b);
/// ```
#[macro_export]
macro_rules! azip {
    (($a:ident in $pa:expr, $b:ident in $pb:expr, $c:ident in $pc:expr) $a2:ident = $b2:ident * $c2:ident + $a3:ident) => {
        $crate::Zip::from($pa).and($pb).and($pc).for_each($crate::Op::MulAddAssign)
    };
    (($a:ident in $pa:expr, $b:ident in $pb:expr, $c:ident in $pc:expr) $a2:ident = $b2:ident + $c2:ident) => { $crate::Zip::from($pa).and($pb).and($pc).for_each($crate::Op::Add) };
    (($a:ident in $pa:expr, $b:ident in $pb:expr, $c:ident in $pc:expr) $a2:ident = $b2:ident - $c2:ident) => { $crate::Zip::from($pa).and($pb).and($pc).for_each($crate::Op::Sub) };
    (($a:ident in $pa:expr, $b:ident in $pb:expr, $c:ident in $pc:expr) $a2:ident = $b2:ident * $c2:ident) => { $crate::Zip::from($pa).and($pb).and($pc).for_each($crate::Op::Mul) };
    (($a:ident in $pa:expr, $b:ident in $pb:expr, $c:ident in $pc:expr) $a2:ident = $b2:ident / $c2:ident) => { $crate::Zip::from($pa).and($pb).and($pc).for_each($crate::Op::Div) };
    (($a:ident in $pa:expr, $b:ident in $pb:expr) $a2:ident = $b2:ident) => { $crate::Zip::from($pa).and($pb).for_each($crate::Op::Assign) };
    (($a:ident in $pa:expr, $b:ident in $pb:expr) $a2:ident += $b2:ident) => { $crate::Zip::from($pa).and($pb).for_each($crate::Op::AddAssign) };
    (($a:ident in $pa:expr, $b:ident in $pb:expr) $a2:ident -= $b2:ident) => { $crate::Zip::from($pa).and($pb).for_each($crate::Op::SubAssign) };
    (($a:ident in $pa:expr, $b:ident in $pb:expr) $a2:ident *= $b2:ident) => { $crate::Zip::from($pa).and($pb).for_each($crate::Op::MulAssign) };
    (($a:ident in $pa:expr, $b:ident in $pb:expr) $a2:ident /= $b2:ident) => { $crate::Zip::from($pa).and($pb).for_each($crate::Op::DivAssign) };
}
/// `par_azip!` (src/parallel/zipmacro.rs:53): the GPU grid is the thread pool, so it is the same expansion.
#[macro_export]
macro_rules! par_azip { ($($t:tt)*) => { $crate::azip!($($t)*) }; }

// ---- multi-GPU through the library's own communicator (csrc/dist.cu): no Python, no MPI requirement -------------------
/// NCCL communicator owned by libndb200.  Rank 0 calls `Comm::unique_id()` and ships the 128 bytes to the other ranks by any
/// means the application has; every rank then calls `Comm::init`.
pub struct Comm { ctx: Arc<Context>, pub rank: usize, pub world: usize }
impl Comm {
    pub fn unique_id() -> [u8; 128] { let mut id = [0u8; 128]; check(unsafe { ffi::ndb_comm_unique_id(id.as_mut_ptr() as *mut c_void) }); id }
    pub fn init(ctx: &Arc<Context>, rank: usize, world: usize, id: &[u8; 128]) -> Self {
        check(unsafe { ffi::ndb_comm_init(ctx.0, rank as i32, world as i32, id.as_ptr() as *const c_void) });
        Comm { ctx: ctx.clone(), rank, world }
    }
    /// `general_mat_mul` with A and C row-sharded and B (valid on rank `src`) replicated.
    pub fn general_mat_mul_rowsharded<A: Elem>(&self, alpha: A, a_local: &DeviceArray<A, Ix2>, b: &DeviceArray<A, Ix2>, beta: A, c_local: &mut DeviceArray<A, Ix2>, src: usize) {
        check(unsafe { ffi::ndb_gemm_rowsharded(self.ctx.0, &alpha as *const A as *const c_void, &a_local.view(), &b.view(), &beta as *const A as *const c_void, &c_local.view(), src as i32) });
    }
    /// `sum_axis` / `mean_axis` / `var_axis` / `std_axis` / `fold_axis(min|max)` along the SHARDED axis over all ranks
    /// (`redop` = `ndb_redop`): fused over peer windows when attached and eligible, NCCL otherwise - the library decides,
    /// identically on every rank.  `None` where ndarray returns `None` (mean of an empty axis).
    pub fn reduce_axis0<A: Elem, D: Dimension + RemoveAxis>(&self, redop: i32, shard: &DeviceArray<A, D>, ddof: f64, n_global: usize) -> Option<DeviceArray<A, D::Smaller>> {
        let out = DeviceArray::<A, D::Smaller>::uninit(&self.ctx, shard.dim.remove_axis(Axis(0)));
        let st = unsafe { ffi::ndb_reduce_axis_sharded(self.ctx.0, redop, 0, &shard.view(), &out.view(), ddof, n_global as i64) };
        if st == ffi::NDB_ERR_EMPTY { return None; }
        check(st);
        Some(out)
    }
    /// `sum()` over all shards (`fold` = Sum / Prod / Min / Max), the same value on every rank.
    pub fn reduce_all<A: Elem + Default, D: Dimension>(&self, shard: &DeviceArray<A, D>, fold: Fold) -> A {
        let mut out = A::default();
        check(unsafe { ffi::ndb_reduce_all_sharded(self.ctx.0, fold as i32, &shard.view(), &mut out as *mut A as *mut c_void) });
        out
    }
}
impl Drop for Comm { fn drop(&mut self) { unsafe { ffi::ndb_comm_destroy(self.ctx.0); } } }

/// A recorded sequence of calls (CUDA graph): `let g = Graph::record(&ctx, || { ...calls... }); g.launch();`
pub struct Graph { ctx: Arc<Context>, exec: *mut c_void }
impl Graph {
    pub fn record(ctx: &Arc<Context>, body: impl FnOnce()) -> Self {
        check(unsafe { ffi::ndb_graph_begin(ctx.0) });
        body();
        let mut exec = std::ptr::null_mut();
        check(unsafe { ffi::ndb_graph_end(ctx.0, &mut exec) });
        Graph { ctx: ctx.clone(), exec }
    }
    pub fn launch(&self) { check(unsafe { ffi::ndb_graph_launch(self.ctx.0, self.exec) }); }
}
impl Drop for Graph { fn drop(&mut self) { unsafe { ffi::ndb_graph_destroy(self.exec); } } }
