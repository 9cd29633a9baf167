//! `panopaea::dec::manifold` (`panopaea/src/dec/manifold.rs:12-143`): the operator traits and `Manifold2d` with its
//! twelve one-line dispatch methods.  The unused `Hodge0/1/2` matrix wrappers and `apply_matrix*` (all
//! `unimplemented!()` in the reference, `dec/grid.rs:193-199, 240-246, 261-267`) are left out: `sparse.rs` is out of scope.

pub trait DecDomain2d<T> {
    type Simplex0;
    type Simplex1;
    type Simplex2;
}

/// `manifold.rs:86-91` without the sparse-matrix forms
pub trait Hodge<T, Simplex> {
    fn apply(&self, dual: &mut Simplex, primal: &Simplex);
    fn apply_inv(&self, primal: &mut Simplex, dual: &Simplex);
}

/// `manifold.rs:135-138`
pub trait DerivativePrimal<T, S, DS> {
    fn apply(&self, d_src: &mut DS, src: &S);
}

/// `manifold.rs:140-143`
pub trait DerivativeDual<T, S, DS> {
    fn apply(&self, d_src: &mut DS, src: &S);
}

/// `manifold.rs:19-84`: same supertraits, same method names, same dispatch.
pub trait Manifold2d<T>:
    DecDomain2d<T>
    + Hodge<T, <Self as DecDomain2d<T>>::Simplex0>
    + Hodge<T, <Self as DecDomain2d<T>>::Simplex1>
    + Hodge<T, <Self as DecDomain2d<T>>::Simplex2>
    + DerivativePrimal<T, <Self as DecDomain2d<T>>::Simplex0, <Self as DecDomain2d<T>>::Simplex1>
    + DerivativePrimal<T, <Self as DecDomain2d<T>>::Simplex1, <Self as DecDomain2d<T>>::Simplex2>
    + DerivativeDual<T, <Self as DecDomain2d<T>>::Simplex2, <Self as DecDomain2d<T>>::Simplex1>
    + DerivativeDual<T, <Self as DecDomain2d<T>>::Simplex1, <Self as DecDomain2d<T>>::Simplex0>
{
    fn num_elem_0(&self) -> usize;
    fn num_elem_1(&self) -> usize;
    fn num_elem_2(&self) -> usize;

    fn new_simplex_0(&self) -> Self::Simplex0;
    fn new_simplex_1(&self) -> Self::Simplex1;
    fn new_simplex_2(&self) -> Self::Simplex2;

    /// primal 0-forms -> primal 1-forms
    fn derivative_0_primal(&self, d_src: &mut Self::Simplex1, src: &Self::Simplex0) {
        <Self as DerivativePrimal<T, Self::Simplex0, Self::Simplex1>>::apply(self, d_src, src)
    }
    /// dual 0-forms (cells) -> dual 1-forms (edges): the gradient, interior edges only
    fn derivative_0_dual(&self, d_src: &mut Self::Simplex1, src: &Self::Simplex2) {
        <Self as DerivativeDual<T, Self::Simplex2, Self::Simplex1>>::apply(self, d_src, src)
    }
    /// primal 1-forms -> primal 2-forms: the divergence stencil
    fn derivative_1_primal(&self, d_src: &mut Self::Simplex2, src: &Self::Simplex1) {
        <Self as DerivativePrimal<T, Self::Simplex1, Self::Simplex2>>::apply(self, d_src, src)
    }
    /// panics, like the reference's `unimplemented!()` (`dec/grid.rs:308-312`)
    fn derivative_1_dual(&self, d_src: &mut Self::Simplex0, src: &Self::Simplex1) {
        <Self as DerivativeDual<T, Self::Simplex1, Self::Simplex0>>::apply(self, d_src, src)
    }

    fn hodge_0_primal(&self, dual: &mut Self::Simplex0, primal: &Self::Simplex0) {
        <Self as Hodge<T, Self::Simplex0>>::apply(self, dual, primal)
    }
    fn hodge_2_dual(&self, primal: &mut Self::Simplex0, dual: &Self::Simplex0) {
        <Self as Hodge<T, Self::Simplex0>>::apply_inv(self, primal, dual)
    }
    fn hodge_1_primal(&self, dual: &mut Self::Simplex1, primal: &Self::Simplex1) {
        <Self as Hodge<T, Self::Simplex1>>::apply(self, dual, primal)
    }
    fn hodge_1_dual(&self, primal: &mut Self::Simplex1, dual: &Self::Simplex1) {
        <Self as Hodge<T, Self::Simplex1>>::apply_inv(self, primal, dual)
    }
    fn hodge_2_primal(&self, dual: &mut Self::Simplex2, primal: &Self::Simplex2) {
        <Self as Hodge<T, Self::Simplex2>>::apply(self, dual, primal)
    }
    fn hodge_0_dual(&self, primal: &mut Self::Simplex2, dual: &Self::Simplex2) {
        <Self as Hodge<T, Self::Simplex2>>::apply_inv(self, primal, dual)
    }
}
