//! `panopaea::dec` (`panopaea/src/dec/mod.rs`): field containers and the operator traits.
pub mod grid;
pub mod manifold;
