//! The part of the `panopaea` crate that `examples/dec_fluid.rs` uses -- `domain::Grid2d`, `dec::grid::{Simplex0,
//! Simplex1, Simplex2}`, `dec::manifold::Manifold2d`, `math::{LinearView, LinearViewReal}`, `pcg` -- with the same names,
//! generics and argument order, the fields living in B200 HBM behind `include/panopaea_b200.h`.
//!
//! NOT COMPILED in this repository's image (no Rust toolchain; see rust/README.md).  The tested hosts over the same ABI
//! are `panopaea_b200/*.py` and `panopaea_b200/host/panopaea.hpp`; this crate follows the latter module by module.
//! Reference citations are `file:line` under the `msiglreith/panopaea` tree.
pub extern crate panopaea_b200_sys as ffi;

pub mod context;
pub mod dec;
pub mod domain;
pub mod fluid;
pub mod math;
pub mod pcg;

pub use context::Context;
