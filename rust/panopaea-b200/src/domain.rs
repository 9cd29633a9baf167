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
