//! `panopaea::dec::grid` (`panopaea/src/dec/grid.rs`): the field containers `Simplex0/1/2<T>` and the operator
//! implementations for `Grid2d`, every operator one kernel of `libpanopaea_b200`.
//!
//! Layouts are the reference's (`dec/grid.rs:10, 37-62, 76`): `Simplex2` `(h, w)` row-major; `Simplex1` ONE flat buffer,
//! `vy (h+1, w)` first, then `vx (h, w+1)`; `Simplex0` `(h+1, w+1)`.  `to_host()` / `upload()` move the flat
//! `view_linear()` slice in one copy.
use std::marker::PhantomData;
use std::ptr;

use crate::context::Context;
use crate::dec::manifold::{DecDomain2d, DerivativeDual, DerivativePrimal, Hodge, Manifold2d};
use crate::domain::Grid2d;
use crate::ffi;
use crate::math::{LinearMut, LinearRef, LinearView, Real};

/// Half-open index rectangle rows `y.0..y.1`, cols `x.0..x.1`: what the example's `for y in a..b { for x in c..d {..}}`
/// loops address (`examples/dec_fluid.rs:48-57, 70-78, 104-112`).
pub fn rect(y: std::ops::Range<usize>, x: std::ops::Range<usize>) -> ffi::pano_rect {
    ffi::pano_rect { y0: y.start as i64, y1: y.end as i64, x0: x.start as i64, x1: x.end as i64 }
}

macro_rules! device_field {
    ($(#[$doc:meta])* $name:ident, $kind:expr) => {
        $(#[$doc])*
        pub struct $name<T> {
            pub(crate) handle: *mut ffi::pano_field,
            dim: (usize, usize),
            _ctx: Context,
            _elem: PhantomData<T>,
        }

        impl<T: Real> $name<T> {
            pub(crate) fn alloc(grid: &Grid2d) -> Self {
                let (h, w) = grid.dim();
                let mut handle = ptr::null_mut();
                ffi::check(unsafe { ffi::pano_field_new(grid.context().raw(), $kind, T::DTYPE, h, w, &mut handle) });
                $name { handle, dim: (h, w), _ctx: grid.context().clone(), _elem: PhantomData }
            }

            /// number of stored values (`view_linear().len()`)
            pub fn len(&self) -> usize {
                let mut n = 0usize;
                ffi::check(unsafe { ffi::pano_field_num_elem($kind, self.dim.0, self.dim.1, &mut n) });
                n
            }

            /// the flat `view_linear()` slice, copied to the host (synchronises)
            pub fn to_host(&self) -> Vec<T> {
                let n = self.len();
                let mut out = vec![T::zero(); n];
                ffi::check(unsafe { ffi::pano_field_download(self.handle, out.as_mut_ptr() as *mut _, n) });
                out
            }

            /// overwrite the field from a flat host slice of `len()` values
            pub fn upload(&mut self, host: &[T]) {
                ffi::check(unsafe { ffi::pano_field_upload(self.handle, host.as_ptr() as *const _, host.len()) });
            }

            pub fn raw(&self) -> *mut ffi::pano_field {
                self.handle
            }
        }

        impl<T> Drop for $name<T> {
            fn drop(&mut self) {
                unsafe {
                    ffi::pano_field_free(self.handle);
                }
            }
        }

        impl<T> LinearView for $name<T> {
            type Elem = T;
            fn view_linear(&self) -> LinearRef<T> {
                LinearRef { handle: self.handle as *const ffi::pano_field, _borrow: PhantomData }
            }
            fn view_linear_mut(&mut self) -> LinearMut<T> {
                LinearMut { handle: self.handle, _borrow: PhantomData }
            }
        }
    };
}

device_field!(
    /// vertex field `(h+1, w+1)` (`dec/grid.rs:10`); not used by dec_fluid
    Simplex0, ffi::PANO_SIMPLEX0);
device_field!(
    /// staggered MAC velocity (`dec/grid.rs:37-40`): `vy (h+1, w)` then `vx (h, w+1)` in one buffer
    Simplex1, ffi::PANO_SIMPLEX1);
device_field!(
    /// cell-centred scalar `(h, w)` (`dec/grid.rs:76`): pressure, density, right-hand side, CG vectors
    Simplex2, ffi::PANO_SIMPLEX2);

impl<T: Real> Simplex0<T> {
    /// array shape `(h+1, w+1)`
    pub fn dim(&self) -> (usize, usize) {
        (self.dim.0 + 1, self.dim.1 + 1)
    }
}

impl<T: Real> Simplex2<T> {
    /// array shape `(h, w)` (the reference derefs to the `Array2`)
    pub fn dim(&self) -> (usize, usize) {
        self.dim
    }

    /// `d[(y, x)] = value` over a rectangle (`examples/dec_fluid.rs:48-57`); an out-of-range rectangle panics like the index would
    pub fn fill_rect(&mut self, r: ffi::pano_rect, value: T) {
        ffi::check(unsafe { ffi::pano_field_fill_rect(self.handle, ffi::PANO_COMP_ALL, r, value.to_f64()) });
    }

    /// O(1) exchange of the device buffers: replaces the copy-back `density.assign(temp)` (`examples/dec_fluid.rs:62`)
    pub fn swap(&mut self, other: &mut Simplex2<T>) {
        ffi::check(unsafe { ffi::pano_field_swap(self.handle, other.handle) });
    }
}

impl<T: Real> Simplex1<T> {
    /// grid dimensions `(h, w)` (`dec/grid.rs:43-45`)
    pub fn dim(&self) -> (usize, usize) {
        self.dim
    }

    /// `split()` (`dec/grid.rs:48-54`) as host copies: `(vy (h+1) x w, vx h x (w+1))`, row-major
    pub fn split_to_host(&self) -> (Vec<T>, Vec<T>) {
        let (h, w) = self.dim;
        let mut flat = self.to_host();
        let vx = flat.split_off((h + 1) * w);
        (flat, vx)
    }

    /// `vy[(y, x)] = value` over a rectangle (`split_mut()` + index loop in the reference)
    pub fn fill_rect_vy(&mut self, r: ffi::pano_rect, value: T) {
        ffi::check(unsafe { ffi::pano_field_fill_rect(self.handle, ffi::PANO_COMP_VY, r, value.to_f64()) });
    }

    /// `vx[(y, x)] = value` over a rectangle
    pub fn fill_rect_vx(&mut self, r: ffi::pano_rect, value: T) {
        ffi::check(unsafe { ffi::pano_field_fill_rect(self.handle, ffi::PANO_COMP_VX, r, value.to_f64()) });
    }

    /// both components at the same `(y, x)` (`examples/dec_fluid.rs:70-78, 104-112`)
    pub fn fill_rect(&mut self, r: ffi::pano_rect, value: T) {
        ffi::check(unsafe { ffi::pano_field_fill_rect(self.handle, ffi::PANO_COMP_ALL, r, value.to_f64()) });
    }

    pub fn swap(&mut self, other: &mut Simplex1<T>) {
        ffi::check(unsafe { ffi::pano_field_swap(self.handle, other.handle) });
    }
}

// ---------------------------------------------------------------------------------------- operators on Grid2d

impl<T: Real> DecDomain2d<T> for Grid2d {
    type Simplex0 = Simplex0<T>;
    type Simplex1 = Simplex1<T>;
    type Simplex2 = Simplex2<T>;
}

/// `Hodge<Simplex0>` (`dec/grid.rs:102-191`), including its corner-addressing quirk (`:109-115`)
impl<T: Real> Hodge<T, Simplex0<T>> for Grid2d {
    fn apply(&self, dual: &mut Simplex0<T>, primal: &Simplex0<T>) {
        ffi::check(unsafe { ffi::pano_hodge_0_primal(dual.handle, primal.handle) });
    }
    fn apply_inv(&self, primal: &mut Simplex0<T>, dual: &Simplex0<T>) {
        ffi::check(unsafe { ffi::pano_hodge_2_dual(primal.handle, dual.handle) });
    }
}

/// `Hodge<Simplex1>` (`dec/grid.rs:202-247`): `apply` negates vx, `apply_inv` negates vy
impl<T: Real> Hodge<T, Simplex1<T>> for Grid2d {
    fn apply(&self, dual: &mut Simplex1<T>, primal: &Simplex1<T>) {
        ffi::check(unsafe { ffi::pano_hodge_1_primal(dual.handle, primal.handle) });
    }
    fn apply_inv(&self, primal: &mut Simplex1<T>, dual: &Simplex1<T>) {
        ffi::check(unsafe { ffi::pano_hodge_1_dual(primal.handle, dual.handle) });
    }
}

/// `Hodge<Simplex2>` (`dec/grid.rs:249-268`): identity copies
impl<T: Real> Hodge<T, Simplex2<T>> for Grid2d {
    fn apply(&self, dual: &mut Simplex2<T>, primal: &Simplex2<T>) {
        ffi::check(unsafe { ffi::pano_hodge_2_primal(dual.handle, primal.handle) });
    }
    fn apply_inv(&self, primal: &mut Simplex2<T>, dual: &Simplex2<T>) {
        ffi::check(unsafe { ffi::pano_hodge_0_dual(primal.handle, dual.handle) });
    }
}

/// `dec/grid.rs:270-289`
impl<T: Real> DerivativePrimal<T, Simplex0<T>, Simplex1<T>> for Grid2d {
    fn apply(&self, edges: &mut Simplex1<T>, vertices: &Simplex0<T>) {
        ffi::check(unsafe { ffi::pano_derivative_0_primal(edges.handle, vertices.handle) });
    }
}

/// `dec/grid.rs:291-306`
impl<T: Real> DerivativePrimal<T, Simplex1<T>, Simplex2<T>> for Grid2d {
    fn apply(&self, faces: &mut Simplex2<T>, edges: &Simplex1<T>) {
        ffi::check(unsafe { ffi::pano_derivative_1_primal(faces.handle, edges.handle) });
    }
}

/// `dec/grid.rs:314-335`: interior edges only, boundary edges are left untouched
impl<T: Real> DerivativeDual<T, Simplex2<T>, Simplex1<T>> for Grid2d {
    fn apply(&self, edges: &mut Simplex1<T>, faces: &Simplex2<T>) {
        ffi::check(unsafe { ffi::pano_derivative_0_dual(edges.handle, faces.handle) });
    }
}

/// `dec/grid.rs:308-312`: `unimplemented!()` -- the library returns `PANO_ERR_UNIMPLEMENTED`, `check` panics
impl<T: Real> DerivativeDual<T, Simplex1<T>, Simplex0<T>> for Grid2d {
    fn apply(&self, vertices: &mut Simplex0<T>, edges: &Simplex1<T>) {
        ffi::check(unsafe { ffi::pano_derivative_1_dual(vertices.handle, edges.handle) });
    }
}

/// `dec/grid.rs:343-371`
impl<T: Real> Manifold2d<T> for Grid2d {
    fn num_elem_0(&self) -> usize {
        (self.dim().0 + 1) * (self.dim().1 + 1)
    }
    fn num_elem_1(&self) -> usize {
        let (h, w) = self.dim();
        (h + 1) * w + h * (w + 1)
    }
    fn num_elem_2(&self) -> usize {
        self.dim().0 * self.dim().1
    }

    fn new_simplex_0(&self) -> Simplex0<T> {
        Simplex0::alloc(self)
    }
    fn new_simplex_1(&self) -> Simplex1<T> {
        Simplex1::alloc(self)
    }
    fn new_simplex_2(&self) -> Simplex2<T> {
        Simplex2::alloc(self)
    }
}
