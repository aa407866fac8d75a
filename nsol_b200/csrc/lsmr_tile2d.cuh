// lsmr_tile2d.cuh -- a whole 2-D LSMR solve on [A; sqrt(alpha) B] as one persistent cooperative launch with TWO grid barriers per
// inner iteration (the two norms of the Golub-Kahan step) instead of four.
//
// lsmr_coopv.cuh runs an inner iteration of a 2-D problem as four dependent phases -- blur pass along the rows of v, forward
// (+ ||u||), blur pass along the rows of u, adjoint (+ ||v||), the vector update riding on the first -- and every phase costs a
// dependent round trip through L2 plus a grid barrier plus the redundant reduction of the partial sums: 29 us per inner iteration
// at 512^2 (BASELINE config 3), 377 us per primal-dual deconvolution iteration (profiles/r2_latency_configs.md).  Here a CTA owns
// a tile of TH x TW pixels and evaluates the WHOLE separable blur on it from one staged copy of the raw tile + R halo rows /
// columns (periodic wrap): pass along the rows in shared memory, pass along x straight into the consumer.  The passes are no
// longer phases of their own, and the vector update of iteration k (h, hbar, x: needs the scalars of the ||v|| step) rides on the
// forward phase of iteration k + 1, which reads the same v:
//
//     [update_k + forward_{k+1}, ||x||^2, ||u||^2]  barrier  [tests_k, beta]  [adjoint, ||v||^2]  barrier  [alpha, rotations]
//
// The stopping test of iteration k is therefore evaluated after forward_{k+1} has overwritten u -- harmless: x is complete, and
// nothing reads u after the solve.  Scalars: the redundant per-CTA recurrences of lsmr_coopv.cuh.  Per element the operators are
// those of the vector kernels (same taps, rows first then x, same expressions); the blur is accumulated in a different grouping
// than the pass kernels -> compared at 1e-11 (tests/test_gpu_r2.py::test_lsmr_tile_solve_matches_other_paths).
// Operator references: nsol/tikhonov_linear_solver.py:258-274 (A_fw / A_bw), nsol/linear_operators.py:60-68 (wrap), :98-106
// (constant); recurrences scipy lsmr.py:328-479.
#pragma once

#define LT_TW 64
#define LT_TH 16

__device__ __forceinline__ int lt_wrap(int q, int n) {
    q %= n;
    return q < 0 ? q + n : q;
}

// blur of the tile (r0, c0) of `src` (periodic): on return s_yb[r][cc] holds the row-blurred value of image row r0 + r, image
// column c0 - R + cc (cc < TW + 2R), and s_raw the raw tile with R (resp. RA = max(R, 1) after the end) halo rows / columns:
// s_raw[rr][cc] = src[r0 - R + rr][c0 - R + cc].  R = 0: s_raw only.
template <typename T, int R>
struct LtSmem {
    static constexpr int RA = R > 0 ? R : 1;
    static constexpr int RAWH = LT_TH + R + RA, RAWW = LT_TW + R + RA;
    static constexpr int YBW = LT_TW + 2 * R;
    T raw[RAWH][RAWW];
    T yb[R > 0 ? LT_TH : 1][R > 0 ? YBW : 1];
};

template <typename T, int R>
__device__ __forceinline__ void lt_stage_blur(LtSmem<T, R> &sm, const TapsR<T, R> &tz, const T *__restrict__ src, int r0, int c0, int nx, int nz) {
    using L = LtSmem<T, R>;
    __syncthreads();                       // the previous tile's readers are done
    {   // all loads of a thread are issued before the first one is stored: one L2 round trip for the whole tile
        constexpr int NL = (L::RAWH * L::RAWW + FAST_TH - 1) / FAST_TH;
        T tmp[NL];
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            const int idx = (int)threadIdx.x + j * FAST_TH;
            const int rr = idx / L::RAWW, cc = idx - rr * L::RAWW;
            const int qr = lt_wrap(r0 - R + rr, nz), qc = lt_wrap(c0 - R + cc, nx);
            tmp[j] = idx < L::RAWH * L::RAWW ? src[(long long)qr * nx + qc] : T(0);
        }
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            const int idx = (int)threadIdx.x + j * FAST_TH;
            const int rr = idx / L::RAWW, cc = idx - rr * L::RAWW;
            if (idx < L::RAWH * L::RAWW) sm.raw[rr][cc] = tmp[j];
        }
    }
    __syncthreads();
    if (R > 0) {
        for (int idx = threadIdx.x; idx < LT_TH * L::YBW; idx += blockDim.x) {
            const int r = idx / L::YBW, cc = idx - r * L::YBW;
            T acc = T(0);
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) acc += tz.t[k] * sm.raw[r + 2 * R - k][cc];      // input row (r0 + r) - (k - R)
            sm.yb[r][cc] = acc;
        }
        __syncthreads();
    }
}

template <typename T, int R>
__device__ __forceinline__ T lt_blur_x(const LtSmem<T, R> &sm, const TapsR<T, R> &tx, int r, int c) {
    if (R == 0) return sm.raw[r][c];
    T acc = T(0);
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) acc += tx.t[k] * sm.yb[r][c + 2 * R - k];               // input column (c0 + c) - (k - R)
    return acc;
}

template <typename T, int R, int VEC_UNUSED>
__global__ void __launch_bounds__(FAST_TH, 2) lsmr_coopt_kernel(const CoopvArgs<T, R> a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ LsmrScalars S;
    __shared__ LtSmem<T, R> sm;
    const FastvGeom<T> &g = a.g;
    const unsigned nb = gridDim.x;
    double *part_u = a.part, *part_v = a.part + nb, *part_x = a.part + 2 * nb;
    const long long n = g.n;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)nb * blockDim.x;
    const double sa_d = a.sa_dev ? *a.sa_dev : a.sqrt_alpha;
    const int nx = g.nx, nz = g.nz;
    const unsigned tiles_x = (unsigned)((nx + LT_TW - 1) / LT_TW), tiles_z = (unsigned)((nz + LT_TH - 1) / LT_TH);
    const unsigned tiles = tiles_x * tiles_z;
    const TapsR<T, R> &tz = a.taps[0], &tx = a.taps[1];      // numpy axis 0 = rows, axis 1 = x
    T *u0 = a.u, *u1 = a.u + n, *u2 = a.u + 2 * n;
    constexpr int PX = LT_TH * LT_TW / FAST_TH;              // pixels of a tile per thread

    // vhat <- (vhat * inv_alpha) * (-beta) + (A^T u0 + sqrt_alpha B^T u1..) * inv_beta, partial ||v||^2 (fastv_adj_body)
    auto adjoint = [&](int first) {
        const T inv_alpha = (T)S.inv_alpha, inv_beta = (T)S.inv_beta, mbeta = (T)(-S.beta), sa = (T)S.sqrt_alpha;
        double acc = 0.0;
        for (unsigned t = blockIdx.x; t < tiles; t += nb) {
            const int r0 = (int)(t / tiles_x) * LT_TH, c0 = (int)(t % tiles_x) * LT_TW;
            // operands of this thread's pixels first: their L2 round trip overlaps the staging and the blur passes of the tile
            T o_u0[PX], o_u1[PX], o_u1l[PX], o_u2[PX], o_u2l[PX], o_v[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const int idx = (int)threadIdx.x + j * FAST_TH;
                const int gr = r0 + idx / LT_TW, gc = c0 + idx % LT_TW;
                const bool ok = gr < nz && gc < nx;
                const long long i = (long long)gr * nx + gc;
                o_u0[j] = (ok && !a.a_blur) ? u0[i] : T(0);
                o_u1[j] = (ok && g.b_op != NSOL_B_NONE) ? u1[i] : T(0);
                o_u1l[j] = (ok && g.b_op == NSOL_B_GRAD && gc > 0) ? u1[i - 1] : T(0);
                o_u2[j] = (ok && g.b_op == NSOL_B_GRAD) ? u2[i] : T(0);
                o_u2l[j] = (ok && g.b_op == NSOL_B_GRAD && gr > 0) ? u2[i - nx] : T(0);
                o_v[j] = (ok && !first) ? a.v[i] : T(0);
            }
            if (a.a_blur) lt_stage_blur<T, R>(sm, tz, u0, r0, c0, nx, nz);
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const int idx = (int)threadIdx.x + j * FAST_TH;
                const int r = idx / LT_TW, c = idx % LT_TW;
                const int gr = r0 + r, gc = c0 + c;
                if (gr >= nz || gc >= nx) continue;
                const long long i = (long long)gr * nx + gc;
                const T hx = a.a_blur ? lt_blur_x<T, R>(sm, tx, r, c) : o_u0[j];
                T rv = hx * inv_beta;
                if (g.b_op == NSOL_B_GRAD) {
                    const T left = gc > 0 ? o_u1l[j] * inv_beta : T(0);
                    T div = g.wx * left + (-g.wx) * (o_u1[j] * inv_beta);
                    const T lo = gr > 0 ? o_u2l[j] * inv_beta : T(0);
                    div = div + (g.wz * lo + (-g.wz) * (o_u2[j] * inv_beta));
                    rv = rv + sa * div;
                } else if (g.b_op == NSOL_B_IDENTITY) {
                    rv = rv + sa * (o_u1[j] * inv_beta);
                }
                const T vn = first ? rv : (o_v[j] * inv_alpha) * mbeta + rv;
                a.v[i] = vn;
                acc += (double)vn * (double)vn;
            }
        }
        acc = block_sum(acc);
        if (threadIdx.x == 0) part_v[blockIdx.x] = acc;
    };

    // ---- u = [b; sqrt_alpha * b_reg], beta = ||u||   (tikhonov_linear_solver.py:226-256; lsmr.py:239-262) ----------
    {
        const T sa = (T)sa_d;
        const long long total = n * (1 + a.rows_b);
        double acc = 0.0;
        for (long long j = gtid; j < total; j += gstride) {
            T w;
            if (j < n) w = a.b[j];
            else w = a.breg ? sa * a.breg[j - n] : T(0);
            a.u[j] = w;
            acc += (double)w * (double)w;
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
    adjoint(1);
    grid.sync();
    {
        const double ss = reduce_partials(part_v, (int)nb);
        if (threadIdx.x == 0) scal_init_alpha(&S, ss);
        __syncthreads();
    }
    {   // h = v, hbar = 0, x = 0   (lsmr.py:277-278, 253)
        const T inv_alpha = (T)S.inv_alpha;
        for (long long j = gtid; j < n; j += gstride) {
            a.h[j] = a.v[j] * inv_alpha;
            a.hbar[j] = T(0);
            a.x[j] = T(0);
        }
    }
    // (h, hbar, x are touched next by the threads of the TILE that owns the pixel -- a different thread than the flat loop above)
    grid.sync();
    // ---- iterations (lsmr.py:328-479) --------------------------------------------------------------------------------
    bool pending = false;          // the vector update + stopping test of the previous iteration are still to be done
    for (int it = 0; it < a.maxiter; ++it) {
        if (S.done && !pending) break;
        {   // update_{k} (if pending) + forward_{k+1}:  u <- (u * inv_beta) * (-alpha) + [A v; sqrt_alpha B v], v = vhat * inv_alpha
            const T inv_alpha = (T)S.inv_alpha, inv_beta = (T)S.inv_beta, malpha = (T)(-S.alpha), sa = (T)S.sqrt_alpha;
            const T c_hbar = (T)S.c_hbar, c_x = (T)S.c_x, c_h = (T)S.c_h;
            double acc = 0.0, accx = 0.0;
            for (unsigned t = blockIdx.x; t < tiles; t += nb) {
                const int r0 = (int)(t / tiles_x) * LT_TH, c0 = (int)(t % tiles_x) * LT_TW;
                // operands of this thread's pixels first: their L2 round trip overlaps the staging and the blur passes of the tile
                T o_u0[PX], o_u1[PX], o_u2[PX], o_h[PX], o_hb[PX], o_x[PX];
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const int idx = (int)threadIdx.x + j * FAST_TH;
                    const int gr = r0 + idx / LT_TW, gc = c0 + idx % LT_TW;
                    const bool ok = gr < nz && gc < nx;
                    const long long i = (long long)gr * nx + gc;
                    o_u0[j] = ok ? u0[i] : T(0);
                    o_u1[j] = (ok && g.b_op != NSOL_B_NONE) ? u1[i] : T(0);
                    o_u2[j] = (ok && g.b_op == NSOL_B_GRAD) ? u2[i] : T(0);
                    o_h[j] = (ok && pending) ? a.h[i] : T(0);
                    o_hb[j] = (ok && pending) ? a.hbar[i] : T(0);
                    o_x[j] = (ok && pending) ? a.x[i] : T(0);
                }
                lt_stage_blur<T, R>(sm, tz, a.v, r0, c0, nx, nz);          // raw vhat tile (+ halo) always: the gradient reads it
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const int idx = (int)threadIdx.x + j * FAST_TH;
                    const int r = idx / LT_TW, c = idx % LT_TW;
                    const int gr = r0 + r, gc = c0 + c;
                    if (gr >= nz || gc >= nx) continue;
                    const long long i = (long long)gr * nx + gc;
                    const T vraw = sm.raw[r + R][c + R];
                    const T hx = a.a_blur ? lt_blur_x<T, R>(sm, tx, r, c) : vraw;
                    const T vc = vraw * inv_alpha;
                    {
                        const T un = (o_u0[j] * inv_beta) * malpha + hx * inv_alpha;
                        u0[i] = un;
                        acc += (double)un * (double)un;
                    }
                    if (g.b_op == NSOL_B_GRAD) {
                        const T right = gc + 1 < nx ? sm.raw[r + R][c + R + 1] * inv_alpha : T(0);
                        const T dkx = g.wx * right + (-g.wx) * vc;
                        const T unx = (o_u1[j] * inv_beta) * malpha + sa * dkx;
                        u1[i] = unx;
                        acc += (double)unx * (double)unx;
                        const T up = gr + 1 < nz ? sm.raw[r + R + 1][c + R] * inv_alpha : T(0);
                        const T dkz = g.wz * up + (-g.wz) * vc;
                        const T unz = (o_u2[j] * inv_beta) * malpha + sa * dkz;
                        u2[i] = unz;
                        acc += (double)unz * (double)unz;
                    } else if (g.b_op == NSOL_B_IDENTITY) {
                        const T un = (o_u1[j] * inv_beta) * malpha + sa * vc;
                        u1[i] = un;
                        acc += (double)un * (double)un;
                    }
                    if (pending) {   // hbar = c_hbar*hbar + h ; x += c_x*hbar ; h = c_h*h + v   (lsmr.py:373-377, 421)
                        const T hbn = o_hb[j] * c_hbar + o_h[j];
                        const T xn = o_x[j] + c_x * hbn;
                        a.hbar[i] = hbn;
                        a.x[i] = xn;
                        a.h[i] = o_h[j] * c_h + vc;
                        accx += (double)xn * (double)xn;
                    }
                }
            }
            acc = block_sum(acc);
            if (threadIdx.x == 0) part_u[blockIdx.x] = acc;
            if (pending) {
                accx = block_sum(accx);
                if (threadIdx.x == 0) part_x[blockIdx.x] = accx;
            }
        }
        grid.sync();
        if (pending) {
            const double ss = reduce_partials(part_x, (int)nb);
            if (threadIdx.x == 0) scal_tests(&S, ss);
            __syncthreads();
            pending = false;
            if (S.done) break;
        }
        {
            const double ss = reduce_partials(part_u, (int)nb);
            if (threadIdx.x == 0) scal_beta(&S, ss);
            __syncthreads();
        }
        adjoint(0);
        grid.sync();
        {
            const double ss = reduce_partials(part_v, (int)nb);
            if (threadIdx.x == 0) scal_alpha(&S, ss);
            __syncthreads();
        }
        pending = true;
    }
    if (pending) {   // the update of the last iteration has no forward phase to ride on
        const T c_hbar = (T)S.c_hbar, c_x = (T)S.c_x, c_h = (T)S.c_h, inv_alpha = (T)S.inv_alpha;
        double accx = 0.0;
        for (long long j = gtid; j < n; j += gstride) {
            const T hv = a.h[j];
            const T hbn = a.hbar[j] * c_hbar + hv;
            const T xn = a.x[j] + c_x * hbn;
            a.hbar[j] = hbn;
            a.x[j] = xn;
            a.h[j] = hv * c_h + a.v[j] * inv_alpha;
            accx += (double)xn * (double)xn;
        }
        accx = block_sum(accx);
        if (threadIdx.x == 0) part_x[blockIdx.x] = accx;
        grid.sync();
        const double ss = reduce_partials(part_x, (int)nb);
        if (threadIdx.x == 0) scal_tests(&S, ss);
        __syncthreads();
    } else {
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.S = S;
    // ---- clip to the bounds (tikhonov_linear_solver.py:156-158) -----------------------------------------------------------
    for (long long j = gtid; j < n; j += gstride) {
        double d = (double)a.x[j];
        d = d < a.lo ? a.lo : (d > a.hi ? a.hi : d);     // np.clip
        a.xout[j] = (T)d;
    }
}
