// dec_fluid.cpp -- examples/dec_fluid.rs, line for line, on top of host/panopaea.hpp.
// Usage: dec_fluid [steps=100] [mode=composed|fused|multigrid|grid3] [out.bin]
//   composed: the reference's own call sequence (one device kernel per Manifold2d call, the generic
//             precond_conjugate_gradient with the Laplacian as a closure)
//   fused   : pano_fluid_step, one call per loop pass
//   grid3   : the same loop on a 64^3 Grid3d (the library's 3-D addition; out.bin receives density, vel, pressure)
// Prints "step i: Iterations k" like the reference prints "Iterations k" (pcg.rs:61), then a checksum.
// The optional output file receives density (h*w), vel (flat Simplex1) and pressure as raw f64.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "panopaea.hpp"

using namespace panopaea;
using dec::Simplex1;
using dec::Simplex2;

int main(int argc, char **argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 100;
    const std::string mode = argc > 2 ? argv[2] : "composed";
    try {
        Context ctx(0);
        if (mode == "grid3") {
            domain::Grid3d g3(ctx, 64, 64, 64);
            DecFluid3 sim(g3, pano_step3_params{0.05, 0.1, 100, PANO_PRECOND_IDENTITY, pano_box{27, 32, 2, 10, 27, 32}, 1.0, 20.0,
                                                pano_box{25, 35, 35, 40, 25, 35}});
            for (int i = 0; i < steps; ++i) printf("step %d: Iterations %ld\n", i, (long)sim.step().iterations);
            if (argc > 3) {
                FILE *f = fopen(argv[3], "wb");
                if (!f) return 2;
                for (const auto &v : {sim.density.to_host(), sim.vel.to_host(), sim.pressure.to_host()}) fwrite(v.data(), 8, v.size(), f);
                fclose(f);
            }
            return 0;
        }
        domain::Grid2d grid(ctx, {128, 128});                                   // :27
        auto vel = grid.new_simplex_1<double>();                                // :29
        auto pressure = grid.new_simplex_2<double>();
        auto density = grid.new_simplex_2<double>();
        auto vel_temp = grid.new_simplex_1<double>();                           // :33
        auto vel_primal_temp = grid.new_simplex_1<double>();
        auto temp = grid.new_simplex_2<double>();
        auto pressure_temp = grid.new_simplex_2<double>();
        auto auxiliary = grid.new_simplex_2<double>();                          // :39
        auto residual = grid.new_simplex_2<double>();
        auto search = grid.new_simplex_2<double>();
        const double timestep = 0.05, threshold = 0.1;                          // :43-44
        const pano_rect inflow{5, 20, 54, 64}, obstacle{70, 80, 50, 70};        // :51-52, :72-73
        const size_t h = 128, w = 128;

        pcg::Multigrid multigrid(ctx, {h, w}, timestep, obstacle);              // used by mode "multigrid" only
        const auto laplacian_closure = [&](Simplex2<double> &laplacian, const Simplex2<double> &p) {   // :100-119
            grid.hodge_2_primal(pressure_temp, p);
            grid.derivative_0_dual(vel_temp, pressure_temp);
            vel_temp.fill_rect(PANO_COMP_ALL, obstacle, 0.0);
            grid.hodge_1_dual(vel_primal_temp, vel_temp);
            grid.derivative_1_primal(laplacian, vel_primal_temp);
            laplacian.scale(timestep);
        };

        for (int i = 0; i < steps; ++i) {                                       // :46
            long iterations = 0;
            if (mode == "fused") {
                pano_step_params p{timestep, threshold, 100, PANO_PRECOND_IDENTITY, inflow, 1.0, 20.0, obstacle};
                pano_pcg_info info;
                check(pano_fluid_step(&p, density.handle(), vel.handle(), pressure.handle(), temp.handle(), vel_temp.handle(),
                                      residual.handle(), auxiliary.handle(), search.handle(), &info));
                iterations = info.iterations;
            } else {
                density.fill_rect(PANO_COMP_ALL, inflow, 1.0);                  // :48-57
                vel.fill_rect(PANO_COMP_VY, inflow, 20.0);
                advect(temp, density, timestep, vel);                           // :59
                advect_mac(vel_temp, vel, timestep, vel);                       // :60
                density.assign(temp);                                           // :62
                vel.assign(vel_temp);                                           // :63
                vel_temp.fill(0.0);                                             // :65
                temp.fill(0.0);                                                 // :66
                grid.hodge_1_dual(vel_temp, vel);                               // :69
                vel_temp.fill_rect(PANO_COMP_ALL, obstacle, 0.0);               // :70-78
                grid.derivative_1_primal(temp, vel_temp);                       // :80
                temp.scale(-1.0);                                               // :81-83
                vel_temp.fill(0.0);                                             // :89
                // :91-119; mode "multigrid" hands the same generic loop another Preconditioner object instead of `&()`
                auto out = mode == "multigrid"
                               ? pcg::precond_conjugate_gradient(multigrid, pressure, temp, 100, threshold, residual, auxiliary, search, laplacian_closure)
                               : pcg::precond_conjugate_gradient(pcg::Identity{}, pressure, temp, 100, threshold, residual, auxiliary, search, laplacian_closure);
                iterations = out.iterations;
                grid.hodge_2_primal(pressure_temp, pressure);                   // :124
                grid.derivative_0_dual(vel_temp, pressure_temp);                // :125
                vel.scaled_add(timestep, vel_temp);                             // :126
                vel.fill_rect(PANO_COMP_VX, pano_rect{0, (int64_t)h, 0, 1}, 0.0);                       // :132-135
                vel.fill_rect(PANO_COMP_VX, pano_rect{0, (int64_t)h, (int64_t)w, (int64_t)w + 1}, 0.0);
                vel.fill_rect(PANO_COMP_VY, pano_rect{0, 1, 0, (int64_t)w}, 0.0);                       // :137-140
                vel.fill_rect(PANO_COMP_VY, pano_rect{(int64_t)h, (int64_t)h + 1, 0, (int64_t)w}, 0.0);
            }
            printf("step %d: Iterations %ld\n", i, iterations);
        }
        auto d = density.view_linear();
        auto v = vel.view_linear();
        auto p = pressure.view_linear();
        double sd = 0, sv = 0, sp = 0;
        for (double x : d) sd += x;
        for (double x : v) sv += x < 0 ? -x : x;
        for (double x : p) sp += x < 0 ? -x : x;
        printf("checksum density %.12e |vel| %.12e |pressure| %.12e\n", sd, sv, sp);
        if (argc > 3) {
            FILE *f = fopen(argv[3], "wb");
            if (!f) return 2;
            fwrite(d.data(), 8, d.size(), f);
            fwrite(v.data(), 8, v.size(), f);
            fwrite(p.data(), 8, p.size(), f);
            fclose(f);
        }
    } catch (const Panic &e) {
        fprintf(stderr, "panic: %s\n", e.what());
        return 1;
    }
    return 0;
}
