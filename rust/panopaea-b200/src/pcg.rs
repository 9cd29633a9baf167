//! `panopaea::pcg` (`panopaea/src/pcg.rs:4-82`): the `Preconditioner` trait, the identity `()`, and
//! `precond_conjugate_gradient` with the reference's generic signature and argument order.  The loop runs on the host
//! and every vector operation is one kernel; the operator `a` is still the caller's closure.
//! `solve_grid_laplacian` is the opt-in fused form: the whole solve in ONE persistent kernel for the operator of
//! `examples/dec_fluid.rs:100-119` (scalars never leave the device).
use std::ptr;

use crate::context::Context;
use crate::dec::grid::Simplex2;
use crate::ffi;
use crate::math::{LinearView, LinearViewReal, Real};

/// `pcg.rs:4-6`
pub trait Preconditioner<L> {
    fn apply(&self, dst: &mut L, src: &L);
}

/// `pcg.rs:8-12`: the identity is a copy
impl<A: Real, L: LinearView<Elem = A>> Preconditioner<L> for () {
    fn apply(&self, dst: &mut L, src: &L) {
        dst.view_linear_mut().assign(&src.view_linear());
    }
}

/// `pcg.rs:14-82`.  Zero initial guess, early-out when `max|b| < threshold` (x stays 0 and the scratch fields are not
/// touched), L-infinity absolute stopping test, at most `max_iterations` operator applications, and the trailing
/// search update when the loop runs out.  Prints what the reference prints.
pub fn precond_conjugate_gradient<L, O, P, T>(
    preconditioner: &P,
    x: &mut L,
    b: &L,
    max_iterations: usize,
    threshold: T,
    residual: &mut L,
    auxiliary: &mut L,
    search: &mut L,
    mut a: O,
) where
    P: Preconditioner<L>,
    T: Real + std::ops::Div<Output = T> + std::ops::Neg<Output = T>,
    L: LinearViewReal<T>,
    O: FnMut(&mut L, &L),
{
    x.view_linear_mut().fill(T::zero()); // :32
    let b_max = b.norm_max();
    if b_max < threshold {
        println!("b start norm {:?}", b_max); // :35-38
        return;
    }
    residual.view_linear_mut().assign(&b.view_linear()); // :40
    preconditioner.apply(auxiliary, residual); // :41
    search.view_linear_mut().assign(&auxiliary.view_linear()); // :42
    let mut sigma = auxiliary.dot_linear(&*residual); // :46

    for i in 0..max_iterations {
        a(auxiliary, search); // :51   z = A s
        let alpha = sigma / auxiliary.dot_linear(&*search); // :53
        x.view_linear_mut().scaled_add(alpha, &search.view_linear()); // :55
        residual.view_linear_mut().scaled_add(-alpha, &auxiliary.view_linear()); // :56
        if residual.norm_max() < threshold {
            println!("Iterations {}", i); // :58-63 (zero-based index)
            break;
        }
        preconditioner.apply(auxiliary, residual); // :65
        let sigma_new = auxiliary.dot_linear(&*residual); // :67
        let beta = sigma_new / sigma; // :68
        search.view_linear_mut().xpby(&auxiliary.view_linear(), beta); // :72-77   s = z + beta s
        sigma = sigma_new; // :79
    }
}

/// Jacobi preconditioner for the dec_fluid operator (an addition behind the reference's trait seam).
pub struct Jacobi {
    pub timestep: f64,
    pub obstacle: ffi::pano_rect,
}

impl Preconditioner<Simplex2<f64>> for Jacobi {
    fn apply(&self, dst: &mut Simplex2<f64>, src: &Simplex2<f64>) {
        ffi::check(unsafe { ffi::pano_jacobi_apply(dst.raw(), src.raw(), self.timestep, self.obstacle) });
    }
}

/// Geometric-multigrid V-cycle for the dec_fluid operator, built once per (grid, timestep, obstacle).
pub struct Multigrid {
    raw: *mut ffi::pano_mg,
    _ctx: Context,
}

impl Multigrid {
    pub fn new(ctx: &Context, dim: (usize, usize), timestep: f64, obstacle: ffi::pano_rect) -> Multigrid {
        let mut raw = ptr::null_mut();
        ffi::check(unsafe { ffi::pano_mg_create(ctx.raw(), dim.0, dim.1, timestep, obstacle, &mut raw) });
        Multigrid { raw, _ctx: ctx.clone() }
    }
}

impl Drop for Multigrid {
    fn drop(&mut self) {
        unsafe {
            ffi::pano_mg_destroy(self.raw);
        }
    }
}

impl Preconditioner<Simplex2<f64>> for Multigrid {
    fn apply(&self, dst: &mut Simplex2<f64>, src: &Simplex2<f64>) {
        ffi::check(unsafe { ffi::pano_mg_apply(self.raw, dst.raw(), src.raw()) });
    }
}

/// What the reference only prints (`pcg.rs:36, 61`).
pub type Outcome = ffi::pano_pcg_info;

/// The fused solve: `precond_conjugate_gradient(&(), x, b, max_iterations, threshold, residual, auxiliary, search, A)` with
/// `A` = the closure of `examples/dec_fluid.rs:100-119` for (`timestep`, `obstacle`), as one persistent kernel.
/// `precond`: `ffi::PANO_PRECOND_IDENTITY` (the reference's `&()`), `_JACOBI` or `_MULTIGRID`.
pub fn solve_grid_laplacian(
    precond: i32,
    x: &mut Simplex2<f64>,
    b: &Simplex2<f64>,
    max_iterations: usize,
    threshold: f64,
    residual: &mut Simplex2<f64>,
    auxiliary: &mut Simplex2<f64>,
    search: &mut Simplex2<f64>,
    timestep: f64,
    obstacle: ffi::pano_rect,
) -> Outcome {
    let mut info = Outcome::default();
    ffi::check(unsafe {
        ffi::pano_pcg_solve(precond, x.raw(), b.raw(), max_iterations as i32, threshold, residual.raw(), auxiliary.raw(),
                            search.raw(), timestep, obstacle, &mut info)
    });
    info
}
