//! The example-local functions of `examples/dec_fluid.rs` (`advect` :173, `advect_mac` :213) with their signatures, the
//! fused passes built from the example's loop body, and the whole loop body as one call.
use crate::context::Context;
use crate::dec::grid::{Simplex1, Simplex2};
use crate::ffi;

/// `pub fn advect(dst, src, timestep, vel)` (`examples/dec_fluid.rs:173-211`); `dst` must not alias `src`
pub fn advect(dst: &mut Simplex2<f64>, src: &Simplex2<f64>, timestep: f64, vel: &Simplex1<f64>) {
    ffi::check(unsafe { ffi::pano_advect(dst.raw(), src.raw(), timestep, vel.raw()) });
}

/// `pub fn advect_mac(dst, src, timestep, vel)` (`examples/dec_fluid.rs:213-291`)
pub fn advect_mac(dst: &mut Simplex1<f64>, src: &Simplex1<f64>, timestep: f64, vel: &Simplex1<f64>) {
    ffi::check(unsafe { ffi::pano_advect_mac(dst.raw(), src.raw(), timestep, vel.raw()) });
}

/// both advections of `:59-60` in one pass over the grid (self-advection)
pub fn advect_all(q_dst: &mut Simplex2<f64>, vel_dst: &mut Simplex1<f64>, q_src: &Simplex2<f64>, vel: &Simplex1<f64>, timestep: f64) {
    ffi::check(unsafe { ffi::pano_advect_all(q_dst.raw(), vel_dst.raw(), q_src.raw(), vel.raw(), timestep) });
}

/// `b = -div(vel)` with the obstacle's edges zeroed (`:69-83`), one pass
pub fn neg_divergence(b: &mut Simplex2<f64>, vel: &Simplex1<f64>, obstacle: ffi::pano_rect) {
    ffi::check(unsafe { ffi::pano_neg_divergence(b.raw(), vel.raw(), obstacle, std::ptr::null_mut()) });
}

/// `z = A(s)`: the Laplacian closure of `:100-119`, one pass
pub fn laplacian_apply(z: &mut Simplex2<f64>, s: &Simplex2<f64>, timestep: f64, obstacle: ffi::pano_rect) {
    ffi::check(unsafe { ffi::pano_laplacian_apply(z.raw(), s.raw(), timestep, obstacle) });
}

/// `vel += dt * d0_dual(p)` on interior edges, then the wall loops (`:124-141`), one pass
pub fn project(vel: &mut Simplex1<f64>, pressure: &Simplex2<f64>, timestep: f64) {
    ffi::check(unsafe { ffi::pano_project(vel.raw(), pressure.raw(), timestep) });
}

/// `util::imgproc::transfer` over the whole field plus the vertical flip of `util::png::export`
/// (`panopaea_utils/src/imgproc.rs:2-5`, `png.rs:12`): `h*w` gray bytes, ready for the RGB8 PNG of `:143-164`
pub fn density_to_u8(density: &Simplex2<f64>, lower: f64, upper: f64) -> Vec<u8> {
    let (h, w) = density.dim();
    let mut out = vec![0u8; h * w];
    ffi::check(unsafe { ffi::pano_density_to_u8(density.raw(), lower, upper, out.as_mut_ptr()) });
    out
}

/// Every literal of `examples/dec_fluid.rs:27, 43-44, 51-54, 72-73, 95`.
pub type StepParams = ffi::pano_step_params;

/// The shipped example's constants for an `n x n` grid, rectangles scaled by `n / 128` (n = 128: the example itself).
pub fn smoke_params(n: usize) -> StepParams {
    let k = (n / 128) as i64;
    StepParams {
        timestep: 0.05,
        threshold: 0.1,
        max_iterations: 100,
        precond: ffi::PANO_PRECOND_IDENTITY,
        inflow: ffi::pano_rect { y0: 5 * k, y1: 20 * k, x0: 54 * k, x1: 64 * k },
        inflow_density: 1.0,
        inflow_vy: 20.0,
        obstacle: ffi::pano_rect { y0: 70 * k, y1: 80 * k, x0: 50 * k, x1: 70 * k },
    }
}

/// One pass of the example's loop body (`:46-141`) on device-resident fields: 6 kernel launches, nothing read back
/// unless `want_info`.  The scratch fields are the ones the example already owns (`:33-41`).
#[allow(clippy::too_many_arguments)]
pub fn fluid_step(
    params: &StepParams,
    density: &mut Simplex2<f64>,
    vel: &mut Simplex1<f64>,
    pressure: &mut Simplex2<f64>,
    temp: &mut Simplex2<f64>,
    vel_temp: &mut Simplex1<f64>,
    residual: &mut Simplex2<f64>,
    auxiliary: &mut Simplex2<f64>,
    search: &mut Simplex2<f64>,
    want_info: bool,
) -> Option<ffi::pano_pcg_info> {
    let mut info = ffi::pano_pcg_info::default();
    let info_ptr = if want_info { &mut info as *mut _ } else { std::ptr::null_mut() };
    ffi::check(unsafe {
        ffi::pano_fluid_step(params, density.raw(), vel.raw(), pressure.raw(), temp.raw(), vel_temp.raw(), residual.raw(),
                             auxiliary.raw(), search.raw(), info_ptr)
    });
    if want_info {
        Some(info)
    } else {
        None
    }
}

/// The same step for fields kept in HOST memory, flat `view_linear()` layout (`density`, `pressure`: `h*w`;
/// `vel`: `(h+1)*w + h*(w+1)`): uploads, steps, downloads.  Pinned buffers (`pano_host_alloc`) make the copies fast.
pub fn fluid_step_host(ctx: &Context, params: &StepParams, dim: (usize, usize), density: &mut [f64], vel: &mut [f64],
                       pressure: &mut [f64]) -> ffi::pano_pcg_info {
    let (h, w) = dim;
    assert_eq!(density.len(), h * w);
    assert_eq!(pressure.len(), h * w);
    assert_eq!(vel.len(), (h + 1) * w + h * (w + 1));
    let mut info = ffi::pano_pcg_info::default();
    ffi::check(unsafe {
        ffi::pano_fluid_step_host(ctx.raw(), params, h, w, density.as_mut_ptr(), vel.as_mut_ptr(), pressure.as_mut_ptr(), &mut info)
    });
    info
}
