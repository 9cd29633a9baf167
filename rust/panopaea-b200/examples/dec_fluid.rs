//! `examples/dec_fluid.rs` of msiglreith/panopaea on a B200: same fields, same constants, same call sequence.
//! Edits against the reference, all forced by the fields living in device memory:
//!   * the index loops over rectangles (`:48-57, :70-78, :104-112, :128-141`) are `fill_rect*` calls;
//!   * `for x in temp.iter_mut() { *x = -*x }` (`:81-83`) and the `* timestep` loop (`:116-118`) are `scale`;
//!   * the PNG dump reads the density through `fluid::density_to_u8` (device-side transfer + flip) instead of
//!     `density[(y, x)]`; the PNG encoder itself (`panopaea_utils`) is unchanged and not repeated here.
//! `--fused` runs the same loop body as ONE call per step (`fluid::fluid_step`: 6 kernels, the solve in one persistent
//! kernel); without it every reference call is one kernel and the CG loop runs on the host, as in the reference.
//! NOT COMPILED in this repository's image (no Rust toolchain).
extern crate panopaea_b200 as panopaea;

use std::time::Instant;

use panopaea::dec::grid::{rect, Simplex1, Simplex2};
use panopaea::dec::manifold::Manifold2d;
use panopaea::domain::Grid2d;
use panopaea::fluid::{self, advect, advect_mac};
use panopaea::math::LinearView;
use panopaea::pcg;

fn grid2d_simplex1(grid: &Grid2d) -> Simplex1<f64> {
    <Grid2d as Manifold2d<f64>>::new_simplex_1(grid)
}

fn grid2d_simplex2(grid: &Grid2d) -> Simplex2<f64> {
    <Grid2d as Manifold2d<f64>>::new_simplex_2(grid)
}

fn main() {
    let fused = std::env::args().any(|a| a == "--fused");
    let grid = Grid2d::new((128, 128));
    let (h, w) = grid.dim();

    let mut vel = grid2d_simplex1(&grid);
    let mut pressure = grid2d_simplex2(&grid);
    let mut density = grid2d_simplex2(&grid);

    let mut vel_temp = grid2d_simplex1(&grid);
    let mut vel_primal_temp = grid2d_simplex1(&grid);
    let mut temp = grid2d_simplex2(&grid);
    let mut pressure_temp = grid2d_simplex2(&grid);

    // conjugate gradient
    let mut auxiliary = grid2d_simplex2(&grid);
    let mut residual = grid2d_simplex2(&grid);
    let mut search = grid2d_simplex2(&grid);

    let timestep = 0.05;
    let threshold = 0.1;
    let params = fluid::smoke_params(128);

    for i in 0..1000 {
        if fused {
            let sw = Instant::now();
            let info = fluid::fluid_step(&params, &mut density, &mut vel, &mut pressure, &mut temp, &mut vel_temp, &mut residual,
                                         &mut auxiliary, &mut search, true).unwrap();
            println!("Iterations {}", info.iterations);
            println!("{} ms", sw.elapsed().as_millis());
        } else {
            // inflow
            density.fill_rect(rect(5..20, 54..64), 1.0);
            vel.fill_rect_vy(rect(5..20, 54..64), 20.0);

            advect(&mut temp, &density, timestep, &vel);
            advect_mac(&mut vel_temp, &vel, timestep, &vel);

            density.view_linear_mut().assign(&temp.view_linear());
            vel.view_linear_mut().assign(&vel_temp.view_linear());

            vel_temp.view_linear_mut().fill(0.0);
            temp.view_linear_mut().fill(0.0);

            // divergence
            grid.hodge_1_dual(&mut vel_temp, &vel);
            vel_temp.fill_rect(rect(70..80, 50..70), 0.0);

            grid.derivative_1_primal(&mut temp, &vel_temp);
            temp.view_linear_mut().scale(-1.0);

            let sw = Instant::now();

            vel_temp.view_linear_mut().fill(0.0);

            pcg::precond_conjugate_gradient(
                &(),
                &mut pressure,
                &temp,
                100,
                threshold,
                &mut residual,
                &mut auxiliary,
                &mut search,
                |laplacian, p| {
                    grid.hodge_2_primal(&mut pressure_temp, p);
                    grid.derivative_0_dual(&mut vel_temp, &pressure_temp);

                    vel_temp.fill_rect(rect(70..80, 50..70), 0.0);

                    grid.hodge_1_dual(&mut vel_primal_temp, &vel_temp);
                    grid.derivative_1_primal(laplacian, &vel_primal_temp);
                    laplacian.view_linear_mut().scale(timestep);
                },
            );

            grid.context().sync();
            println!("{} ms", sw.elapsed().as_millis());

            // project velocity
            grid.hodge_2_primal(&mut pressure_temp, &pressure);
            grid.derivative_0_dual(&mut vel_temp, &pressure_temp);
            vel.view_linear_mut().scaled_add(timestep, &vel_temp.view_linear());

            vel.fill_rect_vx(rect(0..h, 0..1), 0.0);
            vel.fill_rect_vx(rect(0..h, w..w + 1), 0.0);
            vel.fill_rect_vy(rect(0..1, 0..w), 0.0);
            vel.fill_rect_vy(rect(h..h + 1, 0..w), 0.0);
        }

        if i % 10 == 0 {
            // gray bytes, already flipped; the reference turns them into RGB8 and calls util::png::export
            let img = fluid::density_to_u8(&density, -2.0, 2.0);
            let lit = img.iter().filter(|&&v| v > 128).count();
            println!("frame {}: {} of {} pixels above mid-gray", i, lit, img.len());
        }
    }
}
