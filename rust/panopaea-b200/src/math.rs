//! `panopaea::math`: `Real`, `LinearView`, `LinearViewReal` (`panopaea/src/math/mod.rs`, `math/linear_view.rs:5-31`).
//!
//! In the reference `view_linear()` aliases a field as a 1-D `ndarray` view and the callers use `fill`, `assign`,
//! `scaled_add` and `dot` on it.  Device memory cannot be handed out as a host view, so the two methods return proxy
//! objects that offer exactly those operations; the call sites read the same
//! (`vel.view_linear_mut().scaled_add(timestep, &vel_temp.view_linear())`, `examples/dec_fluid.rs:126`).
use std::marker::PhantomData;

use crate::ffi;

/// Scalar types of the fields.  Scalars cross the C ABI as `f64`.
pub trait Real: Copy + PartialOrd + std::fmt::Debug + 'static {
    const DTYPE: i32;
    fn zero() -> Self;
    fn to_f64(self) -> f64;
    fn from_f64(v: f64) -> Self;
}

impl Real for f64 {
    const DTYPE: i32 = ffi::PANO_F64;
    fn zero() -> f64 {
        0.0
    }
    fn to_f64(self) -> f64 {
        self
    }
    fn from_f64(v: f64) -> f64 {
        v
    }
}

impl Real for f32 {
    const DTYPE: i32 = ffi::PANO_F32;
    fn zero() -> f32 {
        0.0
    }
    fn to_f64(self) -> f64 {
        self as f64
    }
    fn from_f64(v: f64) -> f32 {
        v as f32
    }
}

/// Read-only flat alias of a device field (`ArrayView<A, Ix1>` in the reference).
pub struct LinearRef<'a, A> {
    pub(crate) handle: *const ffi::pano_field,
    pub(crate) _borrow: PhantomData<&'a A>,
}

/// Mutable flat alias of a device field (`ArrayViewMut<A, Ix1>` in the reference).
pub struct LinearMut<'a, A> {
    pub(crate) handle: *mut ffi::pano_field,
    pub(crate) _borrow: PhantomData<&'a mut A>,
}

impl<'a, A: Real> LinearRef<'a, A> {
    /// `a.dot(&b)` of ndarray (`math/linear_view.rs:13`); shapes must agree, else panic
    pub fn dot(&self, rhs: &LinearRef<A>) -> A {
        let mut out = 0.0f64;
        ffi::check(unsafe { ffi::pano_field_dot(self.handle, rhs.handle, &mut out) });
        A::from_f64(out)
    }
}

impl<'a, A: Real> LinearMut<'a, A> {
    /// ndarray `fill` (`examples/dec_fluid.rs:65-66, 89`; `pcg.rs:32`)
    pub fn fill(&mut self, value: A) {
        ffi::check(unsafe { ffi::pano_field_fill(self.handle, value.to_f64()) });
    }

    /// ndarray `assign` (`examples/dec_fluid.rs:62-63`; `pcg.rs:10, 40, 42`)
    pub fn assign(&mut self, src: &LinearRef<A>) {
        ffi::check(unsafe { ffi::pano_field_assign(self.handle, src.handle) });
    }

    /// ndarray `scaled_add`: `self += alpha * x` (`pcg.rs:55-56`; `examples/dec_fluid.rs:126`)
    pub fn scaled_add(&mut self, alpha: A, x: &LinearRef<A>) {
        ffi::check(unsafe { ffi::pano_field_scaled_add(self.handle, alpha.to_f64(), x.handle) });
    }

    /// `self = a + beta * self`: the indexed search-update loop of `pcg.rs:72-77` as one pass
    pub fn xpby(&mut self, a: &LinearRef<A>, beta: A) {
        ffi::check(unsafe { ffi::pano_field_xpby(self.handle, a.handle, beta.to_f64()) });
    }

    /// `for x in iter_mut() { *x = *x * alpha }` (`examples/dec_fluid.rs:81-83, 116-118`)
    pub fn scale(&mut self, alpha: A) {
        ffi::check(unsafe { ffi::pano_field_scale(self.handle, alpha.to_f64()) });
    }
}

/// `math::LinearView` (`math/linear_view.rs:5-10`)
pub trait LinearView {
    type Elem;
    fn view_linear(&self) -> LinearRef<Self::Elem>;
    fn view_linear_mut(&mut self) -> LinearMut<Self::Elem>;
}

/// `math::LinearViewReal` (`math/linear_view.rs:12-31`): blanket-implemented for every `LinearView`, as in the reference.
pub trait LinearViewReal<A: Real>: LinearView<Elem = A> {
    fn dot_linear<Rhs: LinearView<Elem = A>>(&self, rhs: &Rhs) -> A {
        self.view_linear().dot(&rhs.view_linear())
    }

    /// `max_k |a[k]|`, starting from zero (`math/linear_view.rs:16-30`)
    fn norm_max(&self) -> A {
        let mut out = 0.0f64;
        ffi::check(unsafe { ffi::pano_field_norm_max(self.view_linear().handle, &mut out) });
        A::from_f64(out)
    }
}

impl<T, A: Real> LinearViewReal<A> for T where T: LinearView<Elem = A> {}

/// `math::trilinear` (`panopaea/src/math/interp.rs:23-36`), argument order kept; evaluated by the library's own expression
/// (the one its 3-D advection kernels use), so host and device agree bit for bit.
#[allow(clippy::too_many_arguments)]
pub fn trilinear(a000: f64, a001: f64, a010: f64, a011: f64, a100: f64, a101: f64, a110: f64, a111: f64, s: f64, t: f64, u: f64) -> f64 {
    unsafe { ffi::pano_trilinear(a000, a001, a010, a011, a100, a101, a110, a111, s, t, u) }
}
