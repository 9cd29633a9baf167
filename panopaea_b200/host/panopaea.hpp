// panopaea.hpp -- C++ host-side mirror of the panopaea crate's interface for the grid fluid step,
// over the C ABI of include/panopaea_b200.h.  Module layout follows the crate:
//   panopaea::domain::Grid2d                      panopaea/src/domain/grid.rs:2-15
//   panopaea::dec::{Simplex0,Simplex1,Simplex2}   panopaea/src/dec/grid.rs:10, 37-62, 76
//   Manifold2d methods on Grid2d                  panopaea/src/dec/manifold.rs:19-84, dec/grid.rs:343-371
//   panopaea::math::LinearView(Real) surface      panopaea/src/math/linear_view.rs:5-31
//   panopaea::pcg                                 panopaea/src/pcg.rs:4-82
// The Rust reference cannot be built in this image; this header is the compiled stand-in for the
// Rust shim of INTEGRATION.md (same names, argument order and panicking behaviour, as exceptions).
#pragma once

#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/panopaea_b200.h"

namespace panopaea {

struct Panic : std::runtime_error {   // the reference panics (ndarray shape checks, unimplemented!())
    int code;
    Panic(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != PANO_OK) throw Panic(rc, std::string("panopaea_b200: ") + pano_last_error());
}

class Context {
  public:
    explicit Context(int device = 0) { check(pano_ctx_create(device, nullptr, &h_)); }
    ~Context() { pano_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    pano_ctx *handle() const { return h_; }
    void sync() const { check(pano_ctx_sync(h_)); }

  private:
    pano_ctx *h_ = nullptr;
};

namespace dec {
template <class T> struct dtype_of;
template <> struct dtype_of<double> { static constexpr int value = PANO_F64; };
template <> struct dtype_of<float> { static constexpr int value = PANO_F32; };

// common part of Simplex0/1/2: a device field + the LinearView / LinearViewReal surface
template <class T, int Kind>
class Field {
  public:
    Field(const Context &ctx, std::pair<size_t, size_t> dim) : dim_(dim) {
        check(pano_field_new(ctx.handle(), Kind, dtype_of<T>::value, dim.first, dim.second, &h_));
    }
    ~Field() { pano_field_free(h_); }
    Field(const Field &) = delete;
    Field &operator=(const Field &) = delete;
    Field(Field &&o) noexcept : h_(o.h_), dim_(o.dim_) { o.h_ = nullptr; }
    pano_field *handle() const { return h_; }
    std::pair<size_t, size_t> dim() const { return dim_; }
    size_t len() const {
        size_t n = 0;
        check(pano_field_num_elem(Kind, dim_.first, dim_.second, &n));
        return n;
    }
    // view_linear(): the flat slice, copied to the host
    std::vector<T> view_linear() const {
        std::vector<T> v(len());
        check(pano_field_download(h_, v.data(), v.size()));
        return v;
    }
    void upload(const std::vector<T> &v) { check(pano_field_upload(h_, v.data(), v.size())); }
    // view_linear_mut().fill / assign / scaled_add
    void fill(T v) { check(pano_field_fill(h_, (double)v)); }
    template <int K2> void assign(const Field<T, K2> &src) { check(pano_field_assign(h_, src.handle())); }
    template <int K2> void scaled_add(T alpha, const Field<T, K2> &rhs) { check(pano_field_scaled_add(h_, (double)alpha, rhs.handle())); }
    void scale(T alpha) { check(pano_field_scale(h_, (double)alpha)); }
    // LinearViewReal
    template <int K2> T dot_linear(const Field<T, K2> &rhs) const {
        double v = 0;
        check(pano_field_dot(h_, rhs.handle(), &v));
        return (T)v;
    }
    T norm_max() const {
        double v = 0;
        check(pano_field_norm_max(h_, &v));
        return (T)v;
    }
    // the example's index loops over rectangles (dec_fluid.rs:48-57, 70-78, 104-112, 128-141)
    void fill_rect(int comp, pano_rect r, T v) { check(pano_field_fill_rect(h_, comp, r, (double)v)); }

  private:
    pano_field *h_ = nullptr;
    std::pair<size_t, size_t> dim_;
};
template <class T> using Simplex0 = Field<T, PANO_SIMPLEX0>;
template <class T> using Simplex1 = Field<T, PANO_SIMPLEX1>;
template <class T> using Simplex2 = Field<T, PANO_SIMPLEX2>;
}  // namespace dec

namespace domain {
// Grid2d + `impl Manifold2d<T> for Grid2d`
class Grid2d {
  public:
    Grid2d(const Context &ctx, std::pair<size_t, size_t> dim) : ctx_(ctx), dim_(dim) {}
    std::pair<size_t, size_t> dim() const { return dim_; }
    size_t num_elem_0() const { return (dim_.first + 1) * (dim_.second + 1); }
    size_t num_elem_1() const { return (dim_.first + 1) * dim_.second + dim_.first * (dim_.second + 1); }
    size_t num_elem_2() const { return dim_.first * dim_.second; }
    template <class T> dec::Simplex0<T> new_simplex_0() const { return dec::Simplex0<T>(ctx_, dim_); }
    template <class T> dec::Simplex1<T> new_simplex_1() const { return dec::Simplex1<T>(ctx_, dim_); }
    template <class T> dec::Simplex2<T> new_simplex_2() const { return dec::Simplex2<T>(ctx_, dim_); }
    // operators: (destination, source), as in dec/manifold.rs:46-83
    template <class T> void derivative_0_primal(dec::Simplex1<T> &d, const dec::Simplex0<T> &s) const { check(pano_derivative_0_primal(d.handle(), s.handle())); }
    template <class T> void derivative_0_dual(dec::Simplex1<T> &d, const dec::Simplex2<T> &s) const { check(pano_derivative_0_dual(d.handle(), s.handle())); }
    template <class T> void derivative_1_primal(dec::Simplex2<T> &d, const dec::Simplex1<T> &s) const { check(pano_derivative_1_primal(d.handle(), s.handle())); }
    template <class T> void derivative_1_dual(dec::Simplex0<T> &d, const dec::Simplex1<T> &s) const { check(pano_derivative_1_dual(d.handle(), s.handle())); }
    template <class T> void hodge_0_primal(dec::Simplex0<T> &dual, const dec::Simplex0<T> &primal) const { check(pano_hodge_0_primal(dual.handle(), primal.handle())); }
    template <class T> void hodge_2_dual(dec::Simplex0<T> &primal, const dec::Simplex0<T> &dual) const { check(pano_hodge_2_dual(primal.handle(), dual.handle())); }
    template <class T> void hodge_1_primal(dec::Simplex1<T> &dual, const dec::Simplex1<T> &primal) const { check(pano_hodge_1_primal(dual.handle(), primal.handle())); }
    template <class T> void hodge_1_dual(dec::Simplex1<T> &primal, const dec::Simplex1<T> &dual) const { check(pano_hodge_1_dual(primal.handle(), dual.handle())); }
    template <class T> void hodge_2_primal(dec::Simplex2<T> &dual, const dec::Simplex2<T> &primal) const { check(pano_hodge_2_primal(dual.handle(), primal.handle())); }
    template <class T> void hodge_0_dual(dec::Simplex2<T> &primal, const dec::Simplex2<T> &dual) const { check(pano_hodge_0_dual(primal.handle(), dual.handle())); }
    const Context &ctx() const { return ctx_; }

  private:
    const Context &ctx_;
    std::pair<size_t, size_t> dim_;
};
}  // namespace domain

namespace pcg {
// impl Preconditioner<L> for ()   (pcg.rs:8-12)
struct Identity {
    template <class L> void apply(L &dst, const L &src) const { dst.assign(src); }
};

// Two more impls of the trait (not in the reference) for the operator of examples/dec_fluid.rs:100-119
struct Jacobi {
    double timestep;
    pano_rect obstacle;
    void apply(dec::Simplex2<double> &dst, const dec::Simplex2<double> &src) const {
        check(pano_jacobi_apply(dst.handle(), src.handle(), timestep, obstacle));
    }
};
class Multigrid {
public:
    Multigrid(const Context &ctx, std::pair<size_t, size_t> dim, double timestep, pano_rect obstacle) {
        check(pano_mg_create(ctx.handle(), dim.first, dim.second, timestep, obstacle, &h_));
    }
    ~Multigrid() { pano_mg_destroy(h_); }
    Multigrid(const Multigrid &) = delete;
    Multigrid &operator=(const Multigrid &) = delete;
    void apply(dec::Simplex2<double> &dst, const dec::Simplex2<double> &src) const { check(pano_mg_apply(h_, dst.handle(), src.handle())); }

private:
    pano_mg *h_ = nullptr;
};

struct Outcome {   // what the reference only prints (pcg.rs:36, 61)
    long iterations;
    long applies;
    double final_residual;
};

// pcg.rs:14-82, generic over the field type L, the preconditioner P and the operator closure O
template <class L, class P, class T, class O>
Outcome precond_conjugate_gradient(const P &preconditioner, L &x, const L &b, size_t max_iterations, T threshold, L &residual,
                                   L &auxiliary, L &search, O a) {
    x.fill((T)0);                                              // :32
    const T bmax = b.norm_max();
    if (bmax < threshold) return Outcome{-1, 0, (double)bmax}; // :35-38
    residual.assign(b);                                        // :40
    preconditioner.apply(auxiliary, residual);                 // :41
    search.assign(auxiliary);                                  // :42
    T sigma = auxiliary.dot_linear(residual);                  // :46
    Outcome out{(long)max_iterations, 0, (double)bmax};
    for (size_t i = 0; i < max_iterations; ++i) {              // :48
        a(auxiliary, search);                                  // :51
        ++out.applies;
        const T alpha = sigma / auxiliary.dot_linear(search);  // :53
        x.scaled_add(alpha, search);                           // :55
        residual.scaled_add(-alpha, auxiliary);                // :56
        const T err = residual.norm_max();                     // :58
        out.final_residual = (double)err;
        if (err < threshold) {                                 // :60-63
            out.iterations = (long)i;
            break;
        }
        preconditioner.apply(auxiliary, residual);             // :65
        const T sigma_new = auxiliary.dot_linear(residual);    // :67
        const T beta = sigma_new / sigma;                      // :68
        check(pano_field_xpby(search.handle(), auxiliary.handle(), (double)beta));   // :72-77
        sigma = sigma_new;                                     // :79
    }
    return out;
}

// the fused path: operator fixed to the closure of examples/dec_fluid.rs:100-119
inline pano_pcg_info solve_grid_laplacian(dec::Simplex2<double> &x, const dec::Simplex2<double> &b, int max_iterations, double threshold,
                                          dec::Simplex2<double> &residual, dec::Simplex2<double> &auxiliary,
                                          dec::Simplex2<double> &search, double timestep, pano_rect obstacle) {
    pano_pcg_info info;
    check(pano_pcg_solve(PANO_PRECOND_IDENTITY, x.handle(), b.handle(), max_iterations, threshold, residual.handle(), auxiliary.handle(),
                         search.handle(), timestep, obstacle, &info));
    return info;
}
}  // namespace pcg

// examples/dec_fluid.rs:173, 213
inline void advect(dec::Simplex2<double> &dst, const dec::Simplex2<double> &src, double timestep, const dec::Simplex1<double> &vel) {
    check(pano_advect(dst.handle(), src.handle(), timestep, vel.handle()));
}
inline void advect_mac(dec::Simplex1<double> &dst, const dec::Simplex1<double> &src, double timestep, const dec::Simplex1<double> &vel) {
    check(pano_advect_mac(dst.handle(), src.handle(), timestep, vel.handle()));
}

// ---- Grid3d (panopaea/src/domain/grid.rs:17-20 is the bare struct; math/interp.rs:23-36 `trilinear`): the library's 3-D addition
namespace domain {
class Grid3d {
  public:
    Grid3d(const Context &ctx, size_t d, size_t h, size_t w) : ctx_(&ctx), d_(d), h_(h), w_(w) {}
    size_t depth() const { return d_; }
    size_t height() const { return h_; }
    size_t width() const { return w_; }
    const Context &context() const { return *ctx_; }

  private:
    const Context *ctx_;
    size_t d_, h_, w_;
};
}  // namespace domain

namespace math {
inline double trilinear(double a000, double a001, double a010, double a011, double a100, double a101, double a110, double a111, double s,
                        double t, double u) {
    return pano_trilinear(a000, a001, a010, a011, a100, a101, a110, a111, s, t, u);
}
}  // namespace math

namespace dec {
template <int Kind>
class Field3 {   // PANO_CELL3 / PANO_FACE3, f64
  public:
    explicit Field3(const domain::Grid3d &g) { check(pano_field3_new(g.context().handle(), Kind, g.depth(), g.height(), g.width(), &h_)); }
    ~Field3() { pano_field_free(h_); }
    Field3(const Field3 &) = delete;
    Field3 &operator=(const Field3 &) = delete;
    pano_field *handle() const { return h_; }
    size_t len() const {
        size_t n = 0;
        check(pano_field_info(h_, nullptr, nullptr, nullptr, nullptr, &n));
        return n;
    }
    void upload(const std::vector<double> &host) { check(pano_field_upload(h_, host.data(), host.size())); }
    std::vector<double> to_host() const {
        std::vector<double> out(len());
        check(pano_field_download(h_, out.data(), out.size()));
        return out;
    }
    void fill_box(pano_box box, double v, int comp = PANO_COMP_ALL) { check(pano_field3_fill_box(h_, comp, box, v)); }

  private:
    pano_field *h_ = nullptr;
};
using Cells3 = Field3<PANO_CELL3>;
using Faces3 = Field3<PANO_FACE3>;
}  // namespace dec

// one pass of the dec_fluid loop body on a Grid3d
struct DecFluid3 {
    pano_step3_params params;
    dec::Faces3 vel, vel_temp;
    dec::Cells3 pressure, density, temp, auxiliary, residual, search;
    DecFluid3(const domain::Grid3d &g, const pano_step3_params &p)
        : params(p), vel(g), vel_temp(g), pressure(g), density(g), temp(g), auxiliary(g), residual(g), search(g) {}
    pano_pcg_info step() {
        pano_pcg_info info;
        check(pano_fluid3_step(&params, density.handle(), vel.handle(), pressure.handle(), temp.handle(), vel_temp.handle(),
                               residual.handle(), auxiliary.handle(), search.handle(), &info));
        return info;
    }
};

}  // namespace panopaea
