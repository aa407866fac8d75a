// lsmr_coopv.cuh -- a whole LSMR solve on [A; sqrt(alpha) B] as ONE persistent cooperative launch built from the
// vector (128-bit, radius-specialised) phase bodies of lsmr_fastv.cuh.
//
// Why: on the images of the reference's configurations (512^2: 2 MB per vector, L2-resident) an inner iteration of the
// multi-kernel path is a chain of eight dependent launches -- blur pass, forward, scalar step, blur pass, adjoint,
// scalar step, update, scalar step -- of ~3.7 us each even when replayed from a CUDA graph: 29 us per inner iteration,
// all of it launch / drain latency (profiles/r1_lsmr_paths.md).  Here the grid stays resident: the phases of an
// iteration are separated by grid.sync() (4 per iteration in 2-D), every CTA loops over its share of the row-mapped
// virtual blocks of a phase, and the scalar recurrences (scipy lsmr.py:344-459) are evaluated redundantly by thread 0
// of EVERY block on a shared-memory copy of LsmrScalars -- all blocks reduce the same per-block partial sums in the
// same order, so the copies stay bit-identical (the scheme of lsmr_coop.cuh, whose generic scalar-indexed phases this
// replaces wherever the vector kernels apply).  Per element the arithmetic is exactly that of the vector kernels; only
// the grouping of the partial sums of the norms differs.
// Right-hand side, start vectors, clip (tikhonov_linear_solver.py:226-256, 156-158; lsmr.py:239-326) are phases of the
// same launch; the weight sqrt(alpha) may come from device memory (graph-free primal-dual deconvolution).
#pragma once

template <typename T, int R>
struct CoopvArgs {
    FastvGeom<T> g;               // b_op already reflects alpha <= EPS (B rows dropped)
    int rows_b;
    int a_blur;                   // 1: A = separable periodic blur of radius R on every axis, 0: identity
    TapsR<T, R> taps[3];          // per numpy axis
    const T *b;
    const T *breg;                // rows of B or NULL (= 0)
    T *u, *v, *h, *hbar, *x, *opbuf, *optmp, *xout;
    double *part;                 // [3][gridDim.x]
    LsmrScalars *S;               // global copy, written by block 0 at the end
    double sqrt_alpha;
    const double *sa_dev;         // if not NULL: sqrt(alpha) is read from here
    double lo, hi;
    int maxiter;
    unsigned vgx;                 // x-chunks of the row-mapped grid: ceil(nx / VEC / FAST_TH)
};

template <typename T, int R, int VEC>
__global__ void __launch_bounds__(FAST_TH) lsmr_coopv_kernel(const CoopvArgs<T, R> a) {
    namespace cg = cooperative_groups;
    using V = Vec<T, VEC>;
    cg::grid_group grid = cg::this_grid();
    __shared__ LsmrScalars S;
    const FastvGeom<T> &g = a.g;
    const unsigned nb = gridDim.x;
    double *part_u = a.part, *part_v = a.part + nb, *part_x = a.part + 2 * nb;
    const long long nvec = g.n / VEC;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)nb * blockDim.x;
    const double sa_d = a.sa_dev ? *a.sa_dev : a.sqrt_alpha;
    const int dim = g.dim;
    const int last_ax = dim - 1;
    const unsigned vgy = (unsigned)g.ny, vgz = (unsigned)g.nz;
    const unsigned rows_total = a.vgx * vgy * vgz;

    // all blur passes except the one along x (fused into the consumer): in -> optmp (-> opbuf); grid.sync after each
    auto blur_front = [&](const T *in) -> const T * {
        const T *src = in;
        if (!a.a_blur) return src;
        for (int ax = 0; ax + 1 < dim; ++ax) {
            T *dst = (src == a.optmp) ? a.opbuf : a.optmp;
            const int kaxis = (dim == 3 && ax == 1) ? 1 : 2;
            const unsigned py = kaxis == 1 ? (vgy + FASTV_ROWS - 1) / FASTV_ROWS : vgy;
            const unsigned pz = kaxis == 2 ? (vgz + FASTV_ROWS - 1) / FASTV_ROWS : vgz;
            const unsigned total = a.vgx * py * pz;
            for (unsigned vb = blockIdx.x; vb < total; vb += nb) {
                const unsigned bx = vb % a.vgx, t = vb / a.vgx;
                fastv_blur_pass_body<T, R, VEC>(g, a.taps[ax], kaxis, src, dst, nullptr, nullptr, bx, t % py, t / py);
            }
            grid.sync();
            src = dst;
        }
        return src;
    };
    auto adjoint = [&](const T *src, int first) {
        double acc = 0.0;
        for (unsigned vb = blockIdx.x; vb < rows_total; vb += nb) {
            const unsigned bx = vb % a.vgx, t = vb / a.vgx;
            fastv_adj_body<T, R, VEC>(g, &S, a.taps[last_ax], src, a.u, a.v, first, nullptr, bx, t % vgy, t / vgy, acc);
        }
        acc = block_sum(acc);
        if (threadIdx.x == 0) part_v[blockIdx.x] = acc;
    };

    // ---- u = [b; sqrt_alpha * b_reg], beta = ||u||   (tikhonov_linear_solver.py:226-256; lsmr.py:239-262) ----------
    {
        const T sa = (T)sa_d;
        const long long total = nvec * (1 + a.rows_b);
        double acc = 0.0;
        for (long long j = gtid; j < total; j += gstride) {
            V w;
            if (j < nvec) w = vec_load<T, VEC>(a.b + j * VEC);
            else if (a.breg) {
                w = vec_load<T, VEC>(a.breg + (j - nvec) * VEC);
#pragma unroll
                for (int e = 0; e < VEC; ++e) w.v[e] = sa * w.v[e];
            } else w = vec_zero<T, VEC>();
            vec_store<T, VEC>(a.u + j * VEC, w);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc += (double)w.v[e] * (double)w.v[e];
        }
        acc = block_sum(acc);
        if (threadIdx.x == 0) part_u[blockIdx.x] = acc;
    }
    grid.sync();
    {
        const double ss = reduce_partials(part_u, (int)nb);
        if (threadIdx.x == 0) scal_init_beta(&S, ss, sa_d, a.maxiter);
        __syncthreads();
    }
    // ---- v = A^T u (cold start), alpha = ||v||   (lsmr.py:264-276) -----------------------------------------------------
    adjoint(blur_front(a.u), 1);
    grid.sync();
    {
        const double ss = reduce_partials(part_v, (int)nb);
        if (threadIdx.x == 0) scal_init_alpha(&S, ss);
        __syncthreads();
    }
    {   // h = v, hbar = 0, x = 0   (lsmr.py:277-278, 253)
        const T inv_alpha = (T)S.inv_alpha;
        const V zero = vec_zero<T, VEC>();
        for (long long j = gtid; j < nvec; j += gstride) {
            V w = vec_load<T, VEC>(a.v + j * VEC);
#pragma unroll
            for (int e = 0; e < VEC; ++e) w.v[e] = w.v[e] * inv_alpha;
            vec_store<T, VEC>(a.h + j * VEC, w);
            vec_store<T, VEC>(a.hbar + j * VEC, zero);
            vec_store<T, VEC>(a.x + j * VEC, zero);
        }
    }
    // ---- iterations (lsmr.py:328-479) --------------------------------------------------------------------------------
    bool pending_tests = false;
    for (int it = 0; it < a.maxiter; ++it) {
        if (S.done && !pending_tests) break;
        // the first blur passes of A v only write scratch: they may run before the previous iteration's stopping test
        const T *src = a.v;
        if (a.a_blur && dim > 1) src = blur_front(a.v);
        else if (pending_tests) grid.sync();
        if (pending_tests) {
            const double ss = reduce_partials(part_x, (int)nb);
            if (threadIdx.x == 0) scal_tests(&S, ss);
            __syncthreads();
            pending_tests = false;
            if (S.done) break;
        }
        {   // u <- (u * inv_beta) * (-alpha) + [A v; sqrt_alpha B v]
            double acc = 0.0;
            for (unsigned vb = blockIdx.x; vb < rows_total; vb += nb) {
                const unsigned bx = vb % a.vgx, t = vb / a.vgx;
                fastv_fwd_body<T, R, VEC>(g, &S, a.taps[last_ax], src, a.v, a.u, nullptr, bx, t % vgy, t / vgy, acc);
            }
            acc = block_sum(acc);
            if (threadIdx.x == 0) part_u[blockIdx.x] = acc;
        }
        grid.sync();
        {
            const double ss = reduce_partials(part_u, (int)nb);
            if (threadIdx.x == 0) scal_beta(&S, ss);
            __syncthreads();
        }
        adjoint(blur_front(a.u), 0);      // v <- (v * inv_alpha) * (-beta) + A^T u0 + sqrt_alpha B^T u1..
        grid.sync();
        {
            const double ss = reduce_partials(part_v, (int)nb);
            if (threadIdx.x == 0) scal_alpha(&S, ss);
            __syncthreads();
        }
        {   // hbar = c_hbar*hbar + h ; x += c_x*hbar ; h = c_h*h + v   (lsmr.py:373-377, 421)
            const T c_hbar = (T)S.c_hbar, c_x = (T)S.c_x, c_h = (T)S.c_h, inv_alpha = (T)S.inv_alpha;
            double acc = 0.0;
            for (long long j = gtid; j < nvec; j += gstride) {
                const V hv = vec_load<T, VEC>(a.h + j * VEC), hb = vec_load<T, VEC>(a.hbar + j * VEC), xv = vec_load<T, VEC>(a.x + j * VEC),
                        vv = vec_load<T, VEC>(a.v + j * VEC);
                V hbn, xn, hn;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    hbn.v[e] = hb.v[e] * c_hbar + hv.v[e];
                    xn.v[e] = xv.v[e] + c_x * hbn.v[e];
                    hn.v[e] = hv.v[e] * c_h + vv.v[e] * inv_alpha;
                    acc += (double)xn.v[e] * (double)xn.v[e];
                }
                vec_store<T, VEC>(a.hbar + j * VEC, hbn);
                vec_store<T, VEC>(a.x + j * VEC, xn);
                vec_store<T, VEC>(a.h + j * VEC, hn);
            }
            acc = block_sum(acc);
            if (threadIdx.x == 0) part_x[blockIdx.x] = acc;
        }
        pending_tests = true;
    }
    grid.sync();   // x complete (and the last partial ||x||^2 visible)
    if (pending_tests) {
        const double ss = reduce_partials(part_x, (int)nb);
        if (threadIdx.x == 0) scal_tests(&S, ss);
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.S = S;
    // ---- clip to the bounds (tikhonov_linear_solver.py:156-158) -----------------------------------------------------------
    for (long long j = gtid; j < nvec; j += gstride) {
        V w = vec_load<T, VEC>(a.x + j * VEC);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            double d = (double)w.v[e];
            d = d < a.lo ? a.lo : (d > a.hi ? a.hi : d);     // np.clip
            w.v[e] = (T)d;
        }
        vec_store<T, VEC>(a.xout + j * VEC, w);
    }
}

// The persistent vector solve applies when the vector kernels do, every blurred axis has the same radius (or A is the
// identity) and the problem is small: <= 2^18 elements by default ("lsmr_path": 0 auto, 4 = whenever possible; 1, 2, 3
// select the other paths).  Measured on B200 (profiles/r2_latency_configs.md): 256^2 23.9 us per inner iteration against
// 25.2 us for the graph-replayed multi-kernel path, 512^2 29.0 vs 28.4, 1024^2 54 vs 50 -- a dependent phase costs ~4.8 us
// whether it ends in a grid.sync() or in a kernel boundary inside a CUDA graph, so the persistent form only wins where the
// launch train cannot be graphed (primal-dual deconvolution: 377 vs 404 us per PD iteration at 512^2) or is very short.
#ifndef NSOL_COOPV_MAX_ELEMENTS
#define NSOL_COOPV_MAX_ELEMENTS (1ll << 18)
#endif
static int coopv_radius(const nsol_lsmr_plan *pl) {
    if (pl->desc.a_op != NSOL_A_BLUR) return 0;
    const int r = pl->desc.radius[0];
    for (int ax = 1; ax < pl->gv.dim; ++ax)
        if (pl->desc.radius[ax] != r) return -1;
    return (r >= 1 && r <= FASTV_MAX_R) ? r : -1;
}

static bool coopv_ok(const nsol_lsmr_plan *pl) {
    if (pl->slab || !fastv_ok(pl) || coopv_radius(pl) < 0) return false;
    if (pl->gv.ny > 65535 || pl->gv.nz > 65535) return false;
    const int path = pl->ctx->lsmr_path;
    if (path == 4) return true;
    if (path != 0) return false;
    return pl->gv.n <= NSOL_COOPV_MAX_ELEMENTS;
}

#include "lsmr_tile2d.cuh"

template <typename T, int R>
static int lsmr_solve_coopv_r(nsol_lsmr_plan *pl, double alpha, const void *b_dev, const void *breg_dev, int maxiter, double lo, double hi,
                              void *x_out, cudaStream_t s, const double *sa_dev) {
    constexpr int VEC = FastvCfg<T>::VEC;
    nsol_ctx *ctx = pl->ctx;
    // 2-D: the tile-fused solve with two grid barriers per inner iteration (lsmr_tile2d.cuh; "lsmr_tile" knob: 2 = never)
    const bool tile = pl->gv.dim == 2 && ctx->lsmr_tile != 2;
    static int per_sm_v[64] = {0}, per_sm_t[64] = {0};   // co-resident CTAs per SM of the two kernels of this instantiation, per device
    int *per_sm = tile ? per_sm_t : per_sm_v;            // (0: unknown, -1: unavailable)
    const int dev = ctx->device & 63;
    if (per_sm[dev] == 0) {
        int coop = 0, v = 0;
        NSOL_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
        if (coop && tile) NSOL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, lsmr_coopt_kernel<T, R, VEC>, FAST_TH, 0));
        else if (coop) NSOL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, lsmr_coopv_kernel<T, R, VEC>, FAST_TH, 0));
        per_sm[dev] = v > 0 ? v : -1;
    }
    if (per_sm[dev] < 0) return NSOL_ESTATE;
    const LsqGeom<T> lg = make_geom<T>(pl);
    CoopvArgs<T, R> a;
    a.g = make_fastv_geom<T>(lg);
    const bool use_b = alpha > 1e-10 && pl->rows_b > 0;
    if (!use_b) a.g.b_op = NSOL_B_NONE;
    a.rows_b = use_b ? pl->rows_b : 0;
    a.a_blur = pl->desc.a_op == NSOL_A_BLUR ? 1 : 0;
    for (int ax = 0; ax < 3; ++ax) a.taps[ax] = lsq_taps_r<T, R>(pl, (a.a_blur && ax < pl->gv.dim) ? ax : -1);
    a.b = (const T *)b_dev;
    a.breg = use_b ? (const T *)breg_dev : nullptr;
    a.u = (T *)pl->u; a.v = (T *)pl->v; a.h = (T *)pl->h; a.hbar = (T *)pl->hbar; a.x = (T *)pl->x;
    a.opbuf = (T *)pl->opbuf; a.optmp = (T *)pl->optmp;
    a.xout = (T *)x_out;
    a.S = pl->S;
    a.sqrt_alpha = use_b ? sqrt(alpha) : 0.0;
    a.sa_dev = use_b ? sa_dev : nullptr;
    a.lo = lo; a.hi = hi;
    a.maxiter = maxiter;
    a.vgx = (unsigned)((pl->gv.nx / VEC + FAST_TH - 1) / FAST_TH);
    // grid: one CTA per row-mapped virtual block up to what is co-resident; grid.sync() cost grows with the block count
    long long want = (long long)a.vgx * pl->gv.ny * pl->gv.nz;
    if (tile) want = (long long)((pl->gv.nx + LT_TW - 1) / LT_TW) * ((pl->gv.nz + LT_TH - 1) / LT_TH);
    // measured (profiles/r2_latency_configs.md): flat from 2 CTAs per SM on, slower below (the phases are bound by the latency
    // of their dependent L2 accesses, which more CTAs overlap -- not by grid.sync())
    int cap_per_sm = per_sm[dev] > 4 ? 4 : per_sm[dev];
    long long cap = (long long)cap_per_sm * ctx->sm_count;
    if (ctx->lsmr_blocks > 0 && ctx->lsmr_blocks < cap) cap = ctx->lsmr_blocks;
    const int blocks = (int)(want < cap ? (want > 0 ? want : 1) : cap);
    if (pl->coopv_blocks < blocks) {
        if (pl->coopv_part) {
            NSOL_CUDA(ctx, cudaStreamSynchronize(s));
            cudaFree(pl->coopv_part);
            pl->coopv_part = nullptr;
        }
        NSOL_CUDA(ctx, cudaMalloc((void **)&pl->coopv_part, sizeof(double) * 3 * (size_t)blocks));
        pl->coopv_blocks = blocks;
    }
    a.part = pl->coopv_part;
    void *params[] = {(void *)&a};
    const void *kernel = tile ? (const void *)lsmr_coopt_kernel<T, R, VEC> : (const void *)lsmr_coopv_kernel<T, R, VEC>;
    NSOL_CUDA(ctx, cudaLaunchCooperativeKernel(kernel, dim3(blocks), dim3(FAST_TH), params, 0, s));
    ctx->launches++;
    return NSOL_OK;
}

// returns NSOL_ESTATE when the cooperative launch is unavailable (the caller falls back to the multi-kernel path)
template <typename T>
static int lsmr_solve_coopv(nsol_lsmr_plan *pl, double alpha, const void *b_dev, const void *breg_dev, int maxiter, double lo, double hi,
                            void *x_out, cudaStream_t s, const double *sa_dev = nullptr) {
    switch (coopv_radius(pl)) {
    case 0: return lsmr_solve_coopv_r<T, 0>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s, sa_dev);
    case 1: return lsmr_solve_coopv_r<T, 1>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s, sa_dev);
    case 2: return lsmr_solve_coopv_r<T, 2>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s, sa_dev);
    case 3: return lsmr_solve_coopv_r<T, 3>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s, sa_dev);
    case 4: return lsmr_solve_coopv_r<T, 4>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s, sa_dev);
    case 5: return lsmr_solve_coopv_r<T, 5>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s, sa_dev);
    case 6: return lsmr_solve_coopv_r<T, 6>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s, sa_dev);
    }
    return NSOL_ESTATE;
}
