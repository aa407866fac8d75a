// lsmr_coop.cuh -- a whole LSMR solve (and optionally a whole ADMM run) as ONE cooperative
// kernel launch.
//
// The multi-kernel path (lsmr_solve_t) needs ~2d+6 launches per inner iteration; on the small
// images of the reference's configurations (512^2 = 2 MB per vector, L2-resident) every one of
// them costs more in launch/drain latency than in work.  Here all blocks are co-resident
// (cudaLaunchCooperativeKernel), the phases of an iteration are separated by grid.sync(), and the
// scalar recurrences are evaluated redundantly by thread 0 of EVERY block on a shared-memory copy
// of LsmrScalars: all blocks reduce the same per-block partial sums in the same order, so their
// copies stay bit-identical and no scalar kernel / extra synchronisation is needed.
// Phases per inner iteration (d = 2): blur pass -> sync -> u update (+ last blur pass fused,
// partial ||u||^2) -> sync -> blur pass -> sync -> v update (partial ||v||^2) -> sync -> h/hbar/x
// update (partial ||x||^2; its stopping test is evaluated after the next sync).  4 grid syncs.
#pragma once
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

template <typename T>
struct CoopArgs {
    LsqGeom<T> g;                 // b_op already reflects alpha <= EPS (B rows dropped)
    int rows_b;
    int a_blur;                   // 1: A = separable periodic blur, 0: identity
    long long np_stride[3];       // numpy axis a: element stride / extent / blur radius / tap offset
    int np_extent[3];
    int radius[3];
    int tap_off[3];
    const T *taps;                // device array, all axes concatenated
    const T *b;
    T *breg;                      // rows of B (input of the solve; rewritten by the ADMM shrink)
    T *u, *v, *h, *hbar, *x, *opbuf, *optmp;
    T *xout;
    double *part;                 // [3][nblocks]
    LsmrScalars *S;               // global copy written by block 0 at the end of every solve
    double sqrt_alpha, lo, hi;
    int maxiter;
    // ADMM (admm_iters > 0): after every solve t = grad(xout) + w, v = shrink(t, ell), w = t - v, breg = v - w
    int admm_iters;
    T ell;
    T *admm_v, *admm_w;
    const T *admm_c;              // the ADMM solver's own b_reg (dim * N) or NULL
};

// Row-wise work distribution for the stencil phases: a work item is a run of up to COOP_RUN
// contiguous x-elements of one (y, z) row, so the (x, y, z) coordinates cost one 32-bit division
// per item instead of two 64-bit divisions per element.  f(i, idx) with idx = {ix, iy, iz}.
template <typename T, typename F>
__device__ __forceinline__ void coop_for_each(const LsqGeom<T> &g, F f) {
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned long long rows = (unsigned long long)g.ny * g.nz;
    // run length: as long as possible while every warp still gets about two items
    int run_len = 256;
    while (run_len > 32 && rows * (unsigned)((g.nx + run_len - 1) / run_len) < 2ull * nwarps) run_len >>= 1;
    const unsigned runs = (unsigned)((g.nx + run_len - 1) / run_len);
    const unsigned long long items = rows * runs;
    for (unsigned long long item = warp; item < items; item += nwarps) {
        const unsigned row = (unsigned)(item / runs), run = (unsigned)(item - (unsigned long long)row * runs);
        int idx[3];
        idx[2] = (int)(row / (unsigned)g.ny);
        idx[1] = (int)(row - (unsigned)idx[2] * (unsigned)g.ny);
        const long long base = (long long)row * g.nx;
        const int x_end = min(g.nx, (int)(run + 1) * run_len);
        for (int ix = (int)run * run_len + lane; ix < x_end; ix += 32) {
            idx[0] = ix;
            f(base + ix, idx);
        }
    }
}

// kernel axis (0 = x, 1 = y, 2 = z) of numpy axis `ax`
__device__ __forceinline__ int coop_kaxis(int dim, int ax) { return ax == dim - 1 ? 0 : ((dim == 3 && ax == 1) ? 1 : 2); }

// periodic 1-D convolution along numpy axis `axis` at element i with coordinates idx
template <typename T>
__device__ __forceinline__ T coop_blur_at(const CoopArgs<T> &a, const T *in, long long i, const int idx[3], int axis) {
    const long long st = a.np_stride[axis];
    const int ext = a.np_extent[axis], r = a.radius[axis];
    const int pos = idx[coop_kaxis(a.g.dim, axis)];
    const T *line = in + (i - (long long)pos * st);
    const T *tp = a.taps + a.tap_off[axis];
    T acc = T(0);
    if (r < ext) {      // wrap with one conditional add/subtract
        for (int k = 0; k <= 2 * r; ++k) {
            int q = pos - (k - r);
            q += q < 0 ? ext : 0;
            q -= q >= ext ? ext : 0;
            acc += tp[k] * line[(long long)q * st];
        }
    } else {            // mask wider than the axis: general modulo
        for (int k = 0; k <= 2 * r; ++k) {
            int q = (pos - (k - r)) % ext;
            if (q < 0) q += ext;
            acc += tp[k] * line[(long long)q * st];
        }
    }
    return acc;
}

// all blur passes except the last one: in -> ... -> returns the array the fused last pass reads
template <typename T>
__device__ const T *coop_blur_front(const CoopArgs<T> &a, cg::grid_group &grid, const T *in) {
    const int dim = a.g.dim;
    const T *src = in;
    for (int ax = 0; ax + 1 < dim; ++ax) {
        T *dst = (src == a.optmp) ? a.opbuf : a.optmp;
        coop_for_each(a.g, [&](long long i, const int idx[3]) { dst[i] = coop_blur_at(a, src, i, idx, ax); });
        grid.sync();
        src = dst;
    }
    return src;
}

template <typename T>
__global__ void __launch_bounds__(LSMR_THREADS) lsmr_coop_kernel(const CoopArgs<T> a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ LsmrScalars S;
    const LsqGeom<T> &g = a.g;
    const long long n = g.n;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    const int nb = gridDim.x;
    double *part_u = a.part, *part_v = a.part + nb, *part_x = a.part + 2 * nb;
    const int last_ax = g.dim - 1;
    const T sa = (T)a.sqrt_alpha;

    const int outer = a.admm_iters > 0 ? a.admm_iters : 1;
    for (int oit = 0; oit < outer; ++oit) {
        // ---- u = [b; sqrt_alpha * b_reg], beta = ||u|| -------------------------------------------
        {
            double acc = 0.0;
            const long long total = n * (1 + a.rows_b);
            for (long long i = gtid; i < total; i += gstride) {
                T v = i < n ? a.b[i] : (a.breg ? sa * a.breg[i - n] : T(0));
                a.u[i] = v;
                acc += (double)v * (double)v;
            }
            acc = block_sum(acc);
            if (threadIdx.x == 0) part_u[blockIdx.x] = acc;
        }
        grid.sync();
        {
            const double ss = reduce_partials(part_u, nb);
            if (threadIdx.x == 0) scal_init_beta(&S, ss, a.sqrt_alpha, a.maxiter);
            __syncthreads();
        }
        // ---- v = A^T u (cold start), alpha = ||v|| ------------------------------------------------
        {
            const T *src = a.a_blur ? coop_blur_front(a, grid, a.u) : a.u;
            const T inv_beta = (T)S.inv_beta;
            double acc = 0.0;
            coop_for_each(g, [&](long long i, const int idx[3]) {
                T r = (a.a_blur ? coop_blur_at(a, src, i, idx, last_ax) : src[i]) * inv_beta;
                if (g.b_op == NSOL_B_GRAD) {
                    T div = T(0);
                    for (int k = 0; k < g.dim; ++k) {
                        const T *uk = a.u + (long long)(1 + k) * n;
                        const T lo = (idx[g.axis[k]] > 0) ? uk[i - g.stride[k]] * inv_beta : T(0);
                        const T dk = g.w[k] * lo + (-g.w[k]) * (uk[i] * inv_beta);
                        div = (k == 0) ? dk : div + dk;
                    }
                    r = r + sa * div;
                } else if (g.b_op == NSOL_B_IDENTITY) {
                    r = r + sa * (a.u[n + i] * inv_beta);
                }
                a.v[i] = r;
                acc += (double)r * (double)r;
            });
            acc = block_sum(acc);
            if (threadIdx.x == 0) part_v[blockIdx.x] = acc;
        }
        grid.sync();
        {
            const double ss = reduce_partials(part_v, nb);
            if (threadIdx.x == 0) scal_init_alpha(&S, ss);
            __syncthreads();
        }
        {
            const T inv_alpha = (T)S.inv_alpha;
            for (long long i = gtid; i < n; i += gstride) {   // h = v, hbar = 0, x = 0
                a.h[i] = a.v[i] * inv_alpha;
                a.hbar[i] = T(0);
                a.x[i] = T(0);
            }
        }
        // ---- iterations ---------------------------------------------------------------------------
        bool pending_tests = false;
        for (int it = 0; it < a.maxiter; ++it) {
            if (S.done && !pending_tests) break;
            // first blur passes of A v (they only write scratch, so they may run before the previous
            // iteration's stopping test is known)
            const T *src = a.v;
            if (a.a_blur && g.dim > 1) {
                src = coop_blur_front(a, grid, a.v);
            } else if (pending_tests) {
                grid.sync();
            }
            if (pending_tests) {
                const double ss = reduce_partials(part_x, nb);
                if (threadIdx.x == 0) scal_tests(&S, ss);
                __syncthreads();
                pending_tests = false;
                if (S.done) break;
            }
            {   // u <- (u * inv_beta) * (-alpha) + [A v; sqrt_alpha B v]
                const T inv_alpha = (T)S.inv_alpha, inv_beta = (T)S.inv_beta, malpha = (T)(-S.alpha);
                double acc = 0.0;
                coop_for_each(g, [&](long long i, const int idx[3]) {
                    const T av = (a.a_blur ? coop_blur_at(a, src, i, idx, last_ax) : src[i]) * inv_alpha;
                    T un = (a.u[i] * inv_beta) * malpha + av;
                    a.u[i] = un;
                    acc += (double)un * (double)un;
                    if (g.b_op == NSOL_B_GRAD) {
                        const T vc = a.v[i] * inv_alpha;
                        for (int k = 0; k < g.dim; ++k) {
                            const T hi = (idx[g.axis[k]] + 1 < g.extent[k]) ? a.v[i + g.stride[k]] * inv_alpha : T(0);
                            const T dk = g.w[k] * hi + (-g.w[k]) * vc;
                            T *uk = a.u + (long long)(1 + k) * n;
                            un = (uk[i] * inv_beta) * malpha + sa * dk;
                            uk[i] = un;
                            acc += (double)un * (double)un;
                        }
                    } else if (g.b_op == NSOL_B_IDENTITY) {
                        T *uk = a.u + n;
                        un = (uk[i] * inv_beta) * malpha + sa * (a.v[i] * inv_alpha);
                        uk[i] = un;
                        acc += (double)un * (double)un;
                    }
                });
                acc = block_sum(acc);
                if (threadIdx.x == 0) part_u[blockIdx.x] = acc;
            }
            grid.sync();
            {
                const double ss = reduce_partials(part_u, nb);
                if (threadIdx.x == 0) scal_beta(&S, ss);
                __syncthreads();
            }
            src = a.u;
            if (a.a_blur && g.dim > 1) src = coop_blur_front(a, grid, a.u);
            {   // v <- (v * inv_alpha) * (-beta) + (A^T u0 + sqrt_alpha B^T u1..)
                const T inv_alpha = (T)S.inv_alpha, inv_beta = (T)S.inv_beta, mbeta = (T)(-S.beta);
                double acc = 0.0;
                coop_for_each(g, [&](long long i, const int idx[3]) {
                    T r = (a.a_blur ? coop_blur_at(a, src, i, idx, last_ax) : src[i]) * inv_beta;
                    if (g.b_op == NSOL_B_GRAD) {
                        T div = T(0);
                        for (int k = 0; k < g.dim; ++k) {
                            const T *uk = a.u + (long long)(1 + k) * n;
                            const T lo = (idx[g.axis[k]] > 0) ? uk[i - g.stride[k]] * inv_beta : T(0);
                            const T dk = g.w[k] * lo + (-g.w[k]) * (uk[i] * inv_beta);
                            div = (k == 0) ? dk : div + dk;
                        }
                        r = r + sa * div;
                    } else if (g.b_op == NSOL_B_IDENTITY) {
                        r = r + sa * (a.u[n + i] * inv_beta);
                    }
                    const T vn = (a.v[i] * inv_alpha) * mbeta + r;
                    // v is also read at neighbouring indices by nobody in this phase: in-place is safe
                    a.v[i] = vn;
                    acc += (double)vn * (double)vn;
                });
                acc = block_sum(acc);
                if (threadIdx.x == 0) part_v[blockIdx.x] = acc;
            }
            grid.sync();
            {
                const double ss = reduce_partials(part_v, nb);
                if (threadIdx.x == 0) scal_alpha(&S, ss);
                __syncthreads();
            }
            {   // hbar = c_hbar*hbar + h ; x += c_x*hbar ; h = c_h*h + v
                const T c_hbar = (T)S.c_hbar, c_x = (T)S.c_x, c_h = (T)S.c_h, inv_alpha = (T)S.inv_alpha;
                double acc = 0.0;
                for (long long i = gtid; i < n; i += gstride) {
                    const T hv = a.h[i];
                    const T hb = a.hbar[i] * c_hbar + hv;
                    a.hbar[i] = hb;
                    const T xv = a.x[i] + c_x * hb;
                    a.x[i] = xv;
                    a.h[i] = hv * c_h + a.v[i] * inv_alpha;
                    acc += (double)xv * (double)xv;
                }
                acc = block_sum(acc);
                if (threadIdx.x == 0) part_x[blockIdx.x] = acc;
            }
            pending_tests = true;
        }
        grid.sync();   // x complete (and the last partial ||x||^2 visible)
        if (pending_tests) {
            const double ss = reduce_partials(part_x, nb);
            if (threadIdx.x == 0) scal_tests(&S, ss);
            __syncthreads();
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) *a.S = S;
        // ---- clip, and (ADMM) the v / w update with the right-hand side of the next solve ------------
        for (long long i = gtid; i < n; i += gstride) {
            double v = (double)a.x[i];
            v = v < a.lo ? a.lo : (v > a.hi ? a.hi : v);
            a.xout[i] = (T)v;
        }
        if (a.admm_iters > 0) {
            grid.sync();
            coop_for_each(g, [&](long long i, const int idx[3]) {
                const T xc = a.xout[i];
                T t[3];
                T ss = T(0);
                for (int k = 0; k < g.dim; ++k) {
                    const T hi = (idx[g.axis[k]] + 1 < g.extent[k]) ? a.xout[i + g.stride[k]] : T(0);
                    T tk = g.w[k] * hi + (-g.w[k]) * xc + a.admm_w[(long long)k * n + i];
                    if (a.admm_c) tk = tk - a.admm_c[(long long)k * n + i];
                    t[k] = tk;
                    ss = (k == 0) ? tk * tk : ss + tk * tk;
                }
                const T nrm = sqrt_t(ss);
                const bool on = nrm > a.ell;
                const T soft = max_t(nrm - a.ell, T(0));
                for (int k = 0; k < g.dim; ++k) {
                    const T vk = on ? soft * t[k] / nrm : T(0);
                    const T wk = t[k] - vk;
                    a.admm_v[(long long)k * n + i] = vk;
                    a.admm_w[(long long)k * n + i] = wk;
                    a.breg[(long long)k * n + i] = a.admm_c ? (vk - wk) + a.admm_c[(long long)k * n + i] : vk - wk;
                }
            });
            grid.sync();
        }
    }
}
