//! `panopaea::domain` (`panopaea/src/domain/grid.rs:2-15`).
use crate::context::Context;

/// `Grid2d { dim: (y, x) }` plus the device context its fields are allocated on.  The reference's `Grid2d` is `Copy`;
/// this one is `Clone` only, because it shares ownership of the context.
#[derive(Clone)]
pub struct Grid2d {
    dim: (usize, usize), // (y, x)
    ctx: Context,
}

impl Grid2d {
    /// `Grid2d::new(dim)` (`domain/grid.rs:8-10`), on the calling thread's default context.
    pub fn new(dim: (usize, usize)) -> Self {
        Grid2d { dim, ctx: Context::default_for_thread() }
    }

    pub fn with_context(dim: (usize, usize), ctx: &Context) -> Self {
        Grid2d { dim, ctx: ctx.clone() }
    }

    /// `(y, x)` (`domain/grid.rs:12-14`)
    pub fn dim(&self) -> (usize, usize) {
        self.dim
    }

    pub fn context(&self) -> &Context {
        &self.ctx
    }
}

/// `Grid3d { dim: (z, y, x) }` (`panopaea/src/domain/grid.rs:17-20`: the struct is all the reference has).  The dec_fluid loop
/// body on it is an addition of the library (`pano_fluid3_step` and friends), reached through `ffi`.
#[derive(Clone)]
pub struct Grid3d {
    dim: (usize, usize, usize), // (z, y, x)
    ctx: Context,
}

impl Grid3d {
    pub fn new(dim: (usize, usize, usize)) -> Self {
        Grid3d { dim, ctx: Context::default_for_thread() }
    }

    pub fn with_context(dim: (usize, usize, usize), ctx: &Context) -> Self {
        Grid3d { dim, ctx: ctx.clone() }
    }

    pub fn dim(&self) -> (usize, usize, usize) {
        self.dim
    }

    pub fn context(&self) -> &Context {
        &self.ctx
    }
}
