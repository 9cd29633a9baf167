//! Raw bindings of `include/panopaea_b200.h`, one item per header item, same order.
//! Every function returns `c_int`: 0 = `PANO_OK`, otherwise a `PANO_ERR_*` code; `pano_last_error()` has the message.
//! NOT COMPILED in this repository's image (no Rust toolchain): kept in sync with the header by
//! `tests/test_abi_symbols.py::test_rust_bindings_match_header`, which compares the two symbol lists.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const PANO_OK: c_int = 0;
pub const PANO_ERR_INVALID: c_int = 1;
pub const PANO_ERR_SHAPE: c_int = 2;
pub const PANO_ERR_CUDA: c_int = 3;
pub const PANO_ERR_UNIMPLEMENTED: c_int = 4;
pub const PANO_ERR_TIMEOUT: c_int = 5;
pub const PANO_ERR_COMM: c_int = 6;

pub const PANO_F64: c_int = 0;
pub const PANO_F32: c_int = 1;
pub const PANO_SIMPLEX0: c_int = 0;
pub const PANO_SIMPLEX1: c_int = 1;
pub const PANO_SIMPLEX2: c_int = 2;
pub const PANO_CELL3: c_int = 3;
pub const PANO_FACE3: c_int = 4;
pub const PANO_COMP_ALL: c_int = 0;
pub const PANO_COMP_VY: c_int = 1;
pub const PANO_COMP_VX: c_int = 2;
pub const PANO_COMP_VZ: c_int = 3;
pub const PANO_PRECOND_IDENTITY: c_int = 0;
pub const PANO_PRECOND_JACOBI: c_int = 1;
pub const PANO_PRECOND_MULTIGRID: c_int = 2;
pub const PANO_STEP_PHASES: usize = 5;
pub const PANO_IPC_HANDLE_BYTES: usize = 64;

#[repr(C)]
pub struct pano_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct pano_field {
    _private: [u8; 0],
}
#[repr(C)]
pub struct pano_mg {
    _private: [u8; 0],
}
#[repr(C)]
pub struct pano_dist {
    _private: [u8; 0],
}

/// half-open index rectangle rows `[y0, y1)` x cols `[x0, x1)`
#[repr(C)]
#[derive(Copy, Clone, Debug, Default, PartialEq)]
pub struct pano_rect {
    pub y0: i64,
    pub y1: i64,
    pub x0: i64,
    pub x1: i64,
}

/// what the reference only prints (`pcg.rs:36, 61`)
#[repr(C)]
#[derive(Copy, Clone, Debug, Default, PartialEq)]
pub struct pano_pcg_info {
    pub iterations: i32,
    pub applies: i32,
    pub final_residual: f64,
    pub rhs_max: f64,
}

/// every literal of `examples/dec_fluid.rs:27, 43-44, 51-54, 72-73, 95`
#[repr(C)]
#[derive(Copy, Clone, Debug, PartialEq)]
pub struct pano_step_params {
    pub timestep: f64,
    pub threshold: f64,
    pub max_iterations: i32,
    pub precond: i32,
    pub inflow: pano_rect,
    pub inflow_density: f64,
    pub inflow_vy: f64,
    pub obstacle: pano_rect,
}

/// half-open index box `[z0,z1) x [y0,y1) x [x0,x1)` of a `Grid3d` (`panopaea/src/domain/grid.rs:17-20`)
#[repr(C)]
#[derive(Copy, Clone, Debug, Default, PartialEq, Eq)]
pub struct pano_box {
    pub z0: i64,
    pub z1: i64,
    pub y0: i64,
    pub y1: i64,
    pub x0: i64,
    pub x1: i64,
}

/// `pano_step_params` on a `Grid3d`
#[repr(C)]
#[derive(Copy, Clone, Debug, PartialEq)]
pub struct pano_step3_params {
    pub timestep: f64,
    pub threshold: f64,
    pub max_iterations: i32,
    pub precond: i32,
    pub inflow: pano_box,
    pub inflow_density: f64,
    pub inflow_vy: f64,
    pub obstacle: pano_box,
}

extern "C" {
    // ------------------------------------------------------------------ library
    pub fn pano_version() -> *const c_char;
    pub fn pano_last_error() -> *const c_char;
    pub fn pano_device_count(count: *mut c_int) -> c_int;

    // ------------------------------------------------------------------ context
    pub fn pano_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut pano_ctx) -> c_int;
    pub fn pano_ctx_destroy(ctx: *mut pano_ctx) -> c_int;
    pub fn pano_ctx_sync(ctx: *mut pano_ctx) -> c_int;
    pub fn pano_ctx_stream(ctx: *mut pano_ctx, stream: *mut *mut c_void) -> c_int;
    pub fn pano_ctx_num_sms(ctx: *mut pano_ctx, n: *mut c_int) -> c_int;
    pub fn pano_ctx_launch_count(ctx: *mut pano_ctx, n: *mut u64) -> c_int;
    pub fn pano_timer_start(ctx: *mut pano_ctx) -> c_int;
    pub fn pano_timer_stop_ms(ctx: *mut pano_ctx, ms: *mut f64) -> c_int;
    pub fn pano_timer_mark(ctx: *mut pano_ctx) -> c_int;
    pub fn pano_timer_marks_ms(ctx: *mut pano_ctx, ms_out: *mut f64, cap: c_int, count: *mut c_int) -> c_int;
    pub fn pano_ctx_step_times(ctx: *mut pano_ctx, ms_out: *mut f64, steps: *mut i64) -> c_int;
    pub fn pano_ctx_cg_profile(ctx: *mut pano_ctx, cycles_out: *mut i64) -> c_int;
    pub fn pano_ctx_cg_profile_ctas(ctx: *mut pano_ctx, cycles_out: *mut i64, n: c_int) -> c_int;
    pub fn pano_ctx_set_option(ctx: *mut pano_ctx, key: *const c_char, value: i64) -> c_int;
    pub fn pano_ctx_get_option(ctx: *mut pano_ctx, key: *const c_char, value: *mut i64) -> c_int;
    pub fn pano_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn pano_host_free(p: *mut c_void) -> c_int;

    // ------------------------------------------------------------------- fields
    pub fn pano_field_new(ctx: *mut pano_ctx, kind: c_int, dtype: c_int, h: usize, w: usize, out: *mut *mut pano_field) -> c_int;
    pub fn pano_field_free(f: *mut pano_field) -> c_int;
    pub fn pano_field_num_elem(kind: c_int, h: usize, w: usize, n: *mut usize) -> c_int;
    pub fn pano_field_info(f: *const pano_field, kind: *mut c_int, dtype: *mut c_int, h: *mut usize, w: *mut usize, n: *mut usize) -> c_int;
    pub fn pano_field_device_ptr(f: *const pano_field, ptr: *mut *mut c_void) -> c_int;
    pub fn pano_field_upload(f: *mut pano_field, host: *const c_void, n_elems: usize) -> c_int;
    pub fn pano_field_download(f: *const pano_field, host: *mut c_void, n_elems: usize) -> c_int;
    pub fn pano_field_fill(f: *mut pano_field, value: f64) -> c_int;
    pub fn pano_field_fill_rect(f: *mut pano_field, comp: c_int, rect: pano_rect, value: f64) -> c_int;
    pub fn pano_field_assign(dst: *mut pano_field, src: *const pano_field) -> c_int;
    pub fn pano_field_swap(a: *mut pano_field, b: *mut pano_field) -> c_int;
    pub fn pano_field_scaled_add(y: *mut pano_field, alpha: f64, x: *const pano_field) -> c_int;
    pub fn pano_field_scale(x: *mut pano_field, alpha: f64) -> c_int;
    pub fn pano_field_xpby(dst: *mut pano_field, a: *const pano_field, beta: f64) -> c_int;
    pub fn pano_field_dot(a: *const pano_field, b: *const pano_field, out: *mut f64) -> c_int;
    pub fn pano_field_norm_max(a: *const pano_field, out: *mut f64) -> c_int;

    // ------------------------------------------- Manifold2d operators, one to one
    pub fn pano_hodge_0_primal(dual: *mut pano_field, primal: *const pano_field) -> c_int;
    pub fn pano_hodge_2_dual(primal: *mut pano_field, dual: *const pano_field) -> c_int;
    pub fn pano_hodge_1_primal(dual: *mut pano_field, primal: *const pano_field) -> c_int;
    pub fn pano_hodge_1_dual(primal: *mut pano_field, dual: *const pano_field) -> c_int;
    pub fn pano_hodge_2_primal(dual: *mut pano_field, primal: *const pano_field) -> c_int;
    pub fn pano_hodge_0_dual(primal: *mut pano_field, dual: *const pano_field) -> c_int;
    pub fn pano_derivative_0_primal(edges: *mut pano_field, vertices: *const pano_field) -> c_int;
    pub fn pano_derivative_1_primal(faces: *mut pano_field, edges: *const pano_field) -> c_int;
    pub fn pano_derivative_0_dual(edges: *mut pano_field, faces: *const pano_field) -> c_int;
    pub fn pano_derivative_1_dual(vertices: *mut pano_field, edges: *const pano_field) -> c_int;

    // ----------------------------------------------------------- fused hot path
    pub fn pano_advect(dst: *mut pano_field, src: *const pano_field, timestep: f64, vel: *const pano_field) -> c_int;
    pub fn pano_advect_mac(dst: *mut pano_field, src: *const pano_field, timestep: f64, vel: *const pano_field) -> c_int;
    pub fn pano_advect_all(q_dst: *mut pano_field, vel_dst: *mut pano_field, q_src: *const pano_field, vel: *const pano_field,
                           timestep: f64) -> c_int;
    pub fn pano_neg_divergence(b: *mut pano_field, vel: *const pano_field, obstacle: pano_rect, rhs_max: *mut f64) -> c_int;
    pub fn pano_laplacian_apply(z: *mut pano_field, s: *const pano_field, timestep: f64, obstacle: pano_rect) -> c_int;
    pub fn pano_project(vel: *mut pano_field, pressure: *const pano_field, timestep: f64) -> c_int;

    // ------------------------------------------------------------------ solver
    pub fn pano_pcg_solve(precond: c_int, x: *mut pano_field, b: *const pano_field, max_iterations: i32, threshold: f64,
                          residual: *mut pano_field, auxiliary: *mut pano_field, search: *mut pano_field, timestep: f64,
                          obstacle: pano_rect, info: *mut pano_pcg_info) -> c_int;

    // ---------------------------------------------------------- preconditioners
    pub fn pano_jacobi_apply(dst: *mut pano_field, src: *const pano_field, timestep: f64, obstacle: pano_rect) -> c_int;
    pub fn pano_mg_create(ctx: *mut pano_ctx, h: usize, w: usize, timestep: f64, obstacle: pano_rect, out: *mut *mut pano_mg) -> c_int;
    pub fn pano_mg_destroy(mg: *mut pano_mg) -> c_int;
    pub fn pano_mg_apply(mg: *mut pano_mg, dst: *mut pano_field, src: *const pano_field) -> c_int;
    pub fn pano_mg_levels(mg: *const pano_mg, levels: *mut c_int, tail_levels: *mut c_int) -> c_int;

    // --------------------------------------------------------------------- step
    pub fn pano_fluid_step(params: *const pano_step_params, density: *mut pano_field, vel: *mut pano_field,
                           pressure: *mut pano_field, temp: *mut pano_field, vel_temp: *mut pano_field,
                           residual: *mut pano_field, auxiliary: *mut pano_field, search: *mut pano_field,
                           info: *mut pano_pcg_info) -> c_int;
    pub fn pano_fluid_step_host(ctx: *mut pano_ctx, params: *const pano_step_params, h: usize, w: usize, density: *mut f64,
                                vel: *mut f64, pressure: *mut f64, info: *mut pano_pcg_info) -> c_int;
    pub fn pano_density_to_u8(density: *const pano_field, lower: f64, upper: f64, host_out: *mut u8) -> c_int;

    // ---------------------------------------------------------------- multi-GPU
    pub fn pano_slab_range(h: usize, rank: c_int, nranks: c_int, y0: *mut usize, y1: *mut usize) -> c_int;
    pub fn pano_dist_create(ctx: *mut pano_ctx, h: usize, w: usize, rank: c_int, nranks: c_int, params: *const pano_step_params,
                            out: *mut *mut pano_dist) -> c_int;
    pub fn pano_dist_destroy(d: *mut pano_dist) -> c_int;
    pub fn pano_dist_window(d: *mut pano_dist, ptr: *mut *mut c_void, bytes: *mut usize) -> c_int;
    pub fn pano_dist_ipc_handle(d: *mut pano_dist, handle_out: *mut c_void) -> c_int;
    pub fn pano_dist_connect(d: *mut pano_dist, kind: c_int, peers: *const c_void) -> c_int;
    pub fn pano_dist_set_max_ctas(d: *mut pano_dist, max_ctas: c_int) -> c_int;
    pub fn pano_dist_upload(d: *mut pano_dist, which: c_int, host_rows: *const f64) -> c_int;
    pub fn pano_dist_download(d: *mut pano_dist, which: c_int, host_rows: *mut f64, rows: *mut usize) -> c_int;
    pub fn pano_dist_step(d: *mut pano_dist) -> c_int;
    pub fn pano_dist_solve(d: *mut pano_dist) -> c_int;
    pub fn pano_dist_sync(d: *mut pano_dist, info: *mut pano_pcg_info) -> c_int;
    pub fn pano_dist_step_host(d: *mut pano_dist, density_rows: *mut f64, vy_rows: *mut f64, vx_rows: *mut f64,
                               info: *mut pano_pcg_info) -> c_int;

    // ------------------------------------------------------------------- Grid3d
    pub fn pano_field3_new(ctx: *mut pano_ctx, kind: c_int, d: usize, h: usize, w: usize, out: *mut *mut pano_field) -> c_int;
    pub fn pano_field3_num_elem(kind: c_int, d: usize, h: usize, w: usize, n: *mut usize) -> c_int;
    pub fn pano_field3_dim(f: *const pano_field, d: *mut usize, h: *mut usize, w: *mut usize) -> c_int;
    pub fn pano_field3_fill_box(f: *mut pano_field, comp: c_int, bx: pano_box, value: f64) -> c_int;
    pub fn pano_trilinear(a000: f64, a001: f64, a010: f64, a011: f64, a100: f64, a101: f64, a110: f64, a111: f64, s: f64, t: f64,
                          u: f64) -> f64;
    pub fn pano_advect3(dst: *mut pano_field, src: *const pano_field, timestep: f64, vel: *const pano_field) -> c_int;
    pub fn pano_advect3_mac(dst: *mut pano_field, src: *const pano_field, timestep: f64, vel: *const pano_field) -> c_int;
    pub fn pano_advect3_all(q_dst: *mut pano_field, vel_dst: *mut pano_field, q_src: *const pano_field, vel: *const pano_field,
                            timestep: f64) -> c_int;
    pub fn pano_neg_divergence3(b: *mut pano_field, vel: *const pano_field, obstacle: pano_box, rhs_max: *mut f64) -> c_int;
    pub fn pano_laplacian3_apply(z: *mut pano_field, s: *const pano_field, timestep: f64, obstacle: pano_box) -> c_int;
    pub fn pano_project3(vel: *mut pano_field, pressure: *const pano_field, timestep: f64) -> c_int;
    pub fn pano_pcg3_solve(precond: c_int, x: *mut pano_field, b: *const pano_field, max_iterations: i32, threshold: f64,
                           residual: *mut pano_field, auxiliary: *mut pano_field, search: *mut pano_field, timestep: f64,
                           obstacle: pano_box, info: *mut pano_pcg_info) -> c_int;
    pub fn pano_fluid3_step(params: *const pano_step3_params, density: *mut pano_field, vel: *mut pano_field,
                            pressure: *mut pano_field, temp: *mut pano_field, vel_temp: *mut pano_field,
                            residual: *mut pano_field, auxiliary: *mut pano_field, search: *mut pano_field,
                            info: *mut pano_pcg_info) -> c_int;
    pub fn pano_fluid3_step_host(ctx: *mut pano_ctx, params: *const pano_step3_params, d: usize, h: usize, w: usize, density: *mut f64,
                                 vel: *mut f64, pressure: *mut f64, info: *mut pano_pcg_info) -> c_int;
}

/// The reference panics on shape mismatches (`ndarray` `Zip`/`assign`) and on `unimplemented!()`; so does the shim.
#[inline]
pub fn check(rc: c_int) {
    if rc != PANO_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(pano_last_error()) }.to_string_lossy().into_owned();
        panic!("panopaea_b200 error {}: {}", rc, msg);
    }
}
