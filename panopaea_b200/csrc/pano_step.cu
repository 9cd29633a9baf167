// pano_step.cu -- one pass of the example's main loop (examples/dec_fluid.rs:46-141) on
// device-resident fields, and the same step for callers whose fields live in host memory.
//
//   reference                                  here
//   :48-57   inflow index loops                2 rectangle fills
//   :59-60   advect + advect_mac               K1 advect_all (one pass)
//   :62-63   copy-back of both fields          buffer swap inside the handles
//   :65-66, :89  clears of the temporaries     not needed: every consumer fully overwrites them
//   :69-83   hodge/box-zero/d1/negate          K3 neg_divergence (+ max|b|, b.b)
//   :91-119  pcg with the Laplacian closure    persistent CG kernel (pano_cg.cu)
//   :124-141 projection + wall loops           K8 project
#include "pano_internal.cuh"

int pano_advect_launch(pano_ctx *ctx, int dtype, void *q_dst, void *vel_dst, const void *q_src, const void *mac_src,
                       const void *vel, size_t h, size_t w, double dt);
int pano_neg_divergence_launch(pano_ctx *ctx, int dtype, void *b, const void *vel, size_t h, size_t w, pano_rect obstacle,
                               bool want_scalars);
int pano_project_launch(pano_ctx *ctx, int dtype, void *vel, const void *p, size_t h, size_t w, double dt);
int pano_cg_solve_raw(pano_ctx *ctx, int dtype, void *x, const void *b, void *r, void *s0, void *s1, size_t h, size_t w,
                      int max_iterations, double threshold, double timestep, pano_rect obstacle, pano_pcg_info *info);

void pano_workspace_free_all(pano_ctx *ctx) {
    for (auto &kv : ctx->workspaces) {
        PanoWorkspace *ws = kv.second;
        pano_field *fs[] = {ws->density, ws->vel, ws->pressure, ws->temp, ws->vel_temp, ws->residual, ws->auxiliary, ws->search};
        for (pano_field *f : fs) pano_field_free(f);
        delete ws;
    }
    ctx->workspaces.clear();
    for (auto &kv : ctx->workspaces3) {
        PanoWorkspace *ws = kv.second;
        pano_field *fs[] = {ws->density, ws->vel, ws->pressure, ws->temp, ws->vel_temp, ws->residual, ws->auxiliary, ws->search};
        for (pano_field *f : fs) pano_field_free(f);
        delete ws;
    }
    ctx->workspaces3.clear();
}

static int get_workspace(pano_ctx *ctx, size_t h, size_t w, PanoWorkspace **out) {
    auto key = std::make_pair(h, w);
    auto it = ctx->workspaces.find(key);
    if (it != ctx->workspaces.end()) {
        *out = it->second;
        return PANO_OK;
    }
    PanoWorkspace *ws = new PanoWorkspace();
    ws->h = h;
    ws->w = w;
    // the workspace enters the cache only when all eight fields exist: a half-built entry would be found by the next
    // call and dereferenced (the header promises that nothing aborts across the FFI boundary)
    struct { pano_field **f; int kind; } want[] = {
        {&ws->density, PANO_SIMPLEX2}, {&ws->vel, PANO_SIMPLEX1}, {&ws->pressure, PANO_SIMPLEX2}, {&ws->temp, PANO_SIMPLEX2},
        {&ws->vel_temp, PANO_SIMPLEX1}, {&ws->residual, PANO_SIMPLEX2}, {&ws->auxiliary, PANO_SIMPLEX2}, {&ws->search, PANO_SIMPLEX2}};
    for (auto &wf : want) {
        const int rc = pano_field_new(ctx, wf.kind, PANO_F64, h, w, wf.f);
        if (rc != PANO_OK) {
            for (auto &g : want) pano_field_free(*g.f);   // null-safe
            delete ws;
            return rc;
        }
    }
    ctx->workspaces[key] = ws;
    *out = ws;
    return PANO_OK;
}

// after_advect (nullable) runs once the new density is final (right after the advection), long before the solve ends:
// the host-buffer entry point uses it to start the density download on a second stream, under the CG kernel.
typedef int (*PanoStepHook)(pano_ctx *ctx, void *user);

static int fluid_step_impl(const pano_step_params *params, pano_field *density, pano_field *vel, pano_field *pressure,
                           pano_field *temp, pano_field *vel_temp, pano_field *residual, pano_field *auxiliary,
                           pano_field *search, pano_pcg_info *info, PanoStepHook after_advect, void *user) {
    if (!params) PANO_FAIL(PANO_ERR_INVALID, "pano_fluid_step: null params");
    const pano_field *s2[] = {density, pressure, temp, residual, auxiliary, search};
    const char *n2[] = {"density", "pressure", "temp", "residual", "auxiliary", "search"};
    for (int i = 0; i < 6; ++i) {
        char nm[64];
        snprintf(nm, sizeof(nm), "pano_fluid_step(%s)", n2[i]);
        PANO_TRY(pano_check_kind(s2[i], PANO_SIMPLEX2, nm));
        PANO_TRY(pano_check_same(density, s2[i], "pano_fluid_step"));
        for (int j = 0; j < i; ++j)
            if (s2[i]->d == s2[j]->d) PANO_FAIL(PANO_ERR_INVALID, "pano_fluid_step: %s aliases %s", n2[i], n2[j]);
    }
    PANO_TRY(pano_check_kind(vel, PANO_SIMPLEX1, "pano_fluid_step(vel)"));
    PANO_TRY(pano_check_kind(vel_temp, PANO_SIMPLEX1, "pano_fluid_step(vel_temp)"));
    PANO_TRY(pano_check_same(vel, vel_temp, "pano_fluid_step"));
    PANO_TRY(pano_check_grid(density, vel, "pano_fluid_step"));
    if (vel->d == vel_temp->d) PANO_FAIL(PANO_ERR_INVALID, "pano_fluid_step: vel aliases vel_temp");
    if (params->precond != PANO_PRECOND_IDENTITY && params->precond != PANO_PRECOND_JACOBI && params->precond != PANO_PRECOND_MULTIGRID)
        PANO_FAIL(PANO_ERR_INVALID, "pano_fluid_step: unknown preconditioner kind %d", params->precond);
    const size_t h = density->h, w = density->w;
    if (h < 2 || w < 2) PANO_FAIL(PANO_ERR_SHAPE, "pano_fluid_step: grid %zux%zu below 2x2", h, w);
    // inflow writes density[(y,x)] and vy[(y,x)]; the obstacle zeroes vy[(y,x)] and vx[(y,x)]
    PANO_TRY(pano_check_rect_within(params->inflow, h, w, "pano_fluid_step(inflow)"));
    PANO_TRY(pano_check_rect_within(params->obstacle, h, w, "pano_fluid_step(obstacle)"));
    pano_ctx *ctx = density->ctx;
    PANO_TRY(pano_activate(ctx));
    const int dt_ = density->dtype;
    const double dt = params->timestep;

    PANO_TRY(pano_phase_mark(ctx, 0));
    // inflow  :48-57
    PANO_TRY(pano_field_fill_rect(density, PANO_COMP_ALL, params->inflow, params->inflow_density));
    PANO_TRY(pano_field_fill_rect(vel, PANO_COMP_VY, params->inflow, params->inflow_vy));
    PANO_TRY(pano_phase_mark(ctx, 1));
    // advect both fields in one pass, then swap instead of copying back  :59-63
    PANO_TRY(pano_advect_launch(ctx, dt_, temp->d, vel_temp->d, density->d, vel->d, vel->d, h, w, dt));
    PANO_TRY(pano_field_swap(density, temp));
    PANO_TRY(pano_field_swap(vel, vel_temp));
    if (after_advect) PANO_TRY(after_advect(ctx, user));
    PANO_TRY(pano_phase_mark(ctx, 2));
    // b = -div  :69-83   (b lives in `temp`, as in the reference)
    PANO_TRY(pano_neg_divergence_launch(ctx, dt_, temp->d, vel->d, h, w, params->obstacle, false));
    PANO_TRY(pano_phase_mark(ctx, 3));
    // pressure solve  :91-119
    pano_pcg_info pinfo;
    const bool host_loop = params->precond != PANO_PRECOND_IDENTITY;
    if (host_loop)   // Jacobi / multigrid: the loop of pcg.rs:32-80 driven from the host (pano_mg.cu)
        PANO_TRY(pano_pcg_precond_raw(ctx, params->precond, pressure, temp, params->max_iterations, params->threshold, residual,
                                      auxiliary, search, dt, params->obstacle, &pinfo));
    else
        PANO_TRY(pano_cg_solve_raw(ctx, dt_, pressure->d, temp->d, residual->d, search->d, auxiliary->d, h, w,
                                   params->max_iterations, params->threshold, dt, params->obstacle, nullptr));
    PANO_TRY(pano_phase_mark(ctx, 4));
    // projection + walls  :124-141
    PANO_TRY(pano_project_launch(ctx, dt_, vel->d, pressure->d, h, w, dt));
    PANO_TRY(pano_phase_mark(ctx, 5));
    if (info && host_loop) {
        *info = pinfo;
        return PANO_OK;
    }
    if (info) {
        PANO_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(PanoCgControl), cudaMemcpyDeviceToHost, ctx->stream));
        PANO_CUDA(cudaStreamSynchronize(ctx->stream));
        PANO_TRY(pano_check_device_error(ctx, "pano_fluid_step"));
        info->iterations = ctx->h_cg->iterations;
        info->applies = ctx->h_cg->applies;
        info->final_residual = ctx->h_cg->final_residual;
        info->rhs_max = ctx->h_cg->rhs_max;
    }
    return PANO_OK;
}

extern "C" {

int pano_fluid_step(const pano_step_params *params, pano_field *density, pano_field *vel, pano_field *pressure,
                    pano_field *temp, pano_field *vel_temp, pano_field *residual, pano_field *auxiliary,
                    pano_field *search, pano_pcg_info *info) {
    return fluid_step_impl(params, density, vel, pressure, temp, vel_temp, residual, auxiliary, search, info, nullptr, nullptr);
}

struct HostStepCopy {
    PanoWorkspace *ws;
    double *density_host;
    size_t bytes;
};

static int start_density_download(pano_ctx *ctx, void *user) {
    HostStepCopy *c = static_cast<HostStepCopy *>(user);
    PANO_CUDA(cudaEventRecord(ctx->ev_advect, ctx->stream));
    PANO_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_advect, 0));
    PANO_CUDA(cudaMemcpyAsync(c->density_host, c->ws->density->d, c->bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    PANO_CUDA(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    return PANO_OK;
}

int pano_fluid_step_host(pano_ctx *ctx, const pano_step_params *params, size_t h, size_t w, double *density, double *vel,
                         double *pressure, pano_pcg_info *info) {
    if (!ctx || !params || !density || !vel) PANO_FAIL(PANO_ERR_INVALID, "pano_fluid_step_host: null argument");
    PANO_TRY(pano_activate(ctx));
    PanoWorkspace *ws = nullptr;
    PANO_TRY(get_workspace(ctx, h, w, &ws));
    const size_t n2 = h * w * sizeof(double), n1 = pano_num_elem(PANO_SIMPLEX1, h, w) * sizeof(double);
    PANO_CUDA(cudaMemcpyAsync(ws->density->d, density, n2, cudaMemcpyHostToDevice, ctx->stream));
    PANO_CUDA(cudaMemcpyAsync(ws->vel->d, vel, n1, cudaMemcpyHostToDevice, ctx->stream));
    HostStepCopy hook{ws, density, n2};
    pano_pcg_info hinfo;
    const bool host_loop = params->precond != PANO_PRECOND_IDENTITY;
    PANO_TRY(fluid_step_impl(params, ws->density, ws->vel, ws->pressure, ws->temp, ws->vel_temp, ws->residual, ws->auxiliary,
                             ws->search, host_loop ? &hinfo : nullptr, start_density_download, &hook));   // density goes home under the solve
    PANO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    PANO_CUDA(cudaMemcpyAsync(vel, ws->vel->d, n1, cudaMemcpyDeviceToHost, ctx->stream));
    if (pressure) PANO_CUDA(cudaMemcpyAsync(pressure, ws->pressure->d, n2, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_loop) {
        PANO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (info) *info = hinfo;
        return PANO_OK;
    }
    PANO_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(PanoCgControl), cudaMemcpyDeviceToHost, ctx->stream));
    PANO_CUDA(cudaStreamSynchronize(ctx->stream));
    PANO_TRY(pano_check_device_error(ctx, "pano_fluid_step_host"));
    if (info) {
        info->iterations = ctx->h_cg->iterations;
        info->applies = ctx->h_cg->applies;
        info->final_residual = ctx->h_cg->final_residual;
        info->rhs_max = ctx->h_cg->rhs_max;
    }
    return PANO_OK;
}

}  // extern "C"
