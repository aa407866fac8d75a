// lsmr_fused2d_v2.cuh -- fused 2-D forward / adjoint kernels of the LSMR iteration, second generation.
//
// Same mapping as the first fused 2-D kernels (lsmr_fastv.cuh: a warp owns a strip of W = 32 * VEC columns and marches
// down a chunk of rows; x-blur through a shared-memory row, register ring of 2R + 1 x-blurred rows along the rows), but
// with the staging of the fused 3-D kernels (lsmr_fused3d.cuh).  ncu on the first generation at 4096^2 float64
// (profiles/r1_lsmr_v2.md): latency-bound -- 96 registers -> 20 warps per SM, input rows fetched through registers one
// row ahead, operand rows one row ahead, long-scoreboard stall 5.5-6.4, issue slots 36-38 % busy, 0.75 of the roofline.
// Here:
//   * every input row of the blurred array (vhat resp. u block 0) is staged with cp.async (LDGSTS), own columns + periodic
//     halo vectors, P rows ahead into a warp-private ring of S = R + P + 1 shared-memory rows; the x-blur reads the
//     staged row directly (no register -> shared-memory copy);
//   * forward: the gradient of row z needs the raw vhat rows z and z + 1 -- still in the ring, so vhat is read ONCE
//     (the first generation fetched them again as operands); the u blocks are staged P rows ahead in an operand ring;
//   * adjoint: u1 (with the vector to its left), u2 and vhat staged the same way, u2 of the previous row in registers;
//   * everything a warp stages is warp-private: one __syncwarp per row, no block barrier.
// DRAM traffic = the algorithmic 7 (forward) / 5 (adjoint) words per pixel plus the 2R / zc warm-up rows of a chunk.
// The blur is accumulated with fused multiply-adds, x first then along the rows; compared with the generic kernels at
// 1e-11 (tests/test_gpu_parity.py::test_lsmr_fused_2d_kernels_match_generic_kernels).
#pragma once

#define F2_WARPS 4
#ifndef F2_P
#define F2_P 3
#endif

template <typename T, int VEC, int R>
struct F2 {
    static constexpr int A = (R + VEC - 1) / VEC * VEC;
    static constexpr int HV = A / VEC;
    static constexpr int W = 32 * VEC;
    static constexpr int LEN = W + 2 * A;
    static constexpr int S = R + F2_P + 1;
    static constexpr int OS = F2_P + 1;
    static constexpr int OPF = 3 * W;                       // forward: u0, u1, u2
    static constexpr int U1LEN = W + VEC;
    static constexpr int OPA = U1LEN + 2 * W;               // adjoint: u1 (+ left vector) | u2 | vhat
    static size_t warp_elems(bool fwd) { return (size_t)S * LEN + (size_t)OS * (fwd ? OPF : OPA); }
    static size_t smem(bool fwd) { return sizeof(T) * F2_WARPS * warp_elems(fwd); }
};

template <typename T, int R, int VEC, bool FWD>
__global__ void __launch_bounds__(32 * F2_WARPS) fused2d_v2_kernel(Fused2dGeom g, T wx, T wz, const LsmrScalars *__restrict__ S, TapsR<T, R> tx,
                                                                  TapsR<T, R> tz, const T *__restrict__ blur_in, T *__restrict__ u,
                                                                  T *__restrict__ vhat, double *__restrict__ part, int first) {
    using V = Vec<T, VEC>;
    using L = F2<T, VEC, R>;
    extern __shared__ __align__(16) unsigned char f2_smem[];
    if (S->done) return;
    constexpr int OPSZ = FWD ? L::OPF : L::OPA;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *s_raw = reinterpret_cast<T *>(f2_smem) + (size_t)warp * (L::S * L::LEN + L::OS * OPSZ);     // [S][LEN]
    T *s_op = s_raw + L::S * L::LEN;                                                                 // [OS][OPSZ]
    const int xw0 = ((int)blockIdx.x * F2_WARPS + warp) * L::W;
    const bool warp_in = xw0 < g.nx;
    const int z0 = (int)blockIdx.y * g.zc;
    const int z1 = min(g.nz, z0 + g.zc);
    const T inv_alpha = (T)S->inv_alpha, inv_beta = (T)S->inv_beta, sa = (T)S->sqrt_alpha;
    const T mscale = FWD ? (T)(-S->alpha) : (T)(-S->beta);
    double acc = 0.0;
    if (warp_in) {
        const int x = xw0 + lane * VEC;
        const bool active = x < g.nx;
        const int wcols = min(L::W, g.nx - xw0);
        const bool has_halo = lane < 2 * L::HV;
        int hcol = lane < L::HV ? xw0 - L::A + lane * VEC : xw0 + wcols + (lane - L::HV) * VEC;
        hcol = f3_mod(hcol, g.nx);
        const int hslot = lane < L::HV ? lane * VEC : L::A + wcols + (lane - L::HV) * VEC;
        T *u0 = u, *u1 = u + g.n, *u2 = u + 2 * g.n;

        auto issue_raw = [&](int j) {        // raw row of ring index j: image row z0 - R + j (periodic)
            int zi = z0 - R + j;
            zi += zi < 0 ? g.nz : 0;
            zi -= zi >= g.nz ? g.nz : 0;
            const T *src = blur_in + (long long)zi * g.nx;
            T *st = s_raw + (j % L::S) * L::LEN;
            if (active) f3_cp16(st + L::A + lane * VEC, src + x);
            if (has_halo) f3_cp16(st + hslot, src + hcol);
        };
        auto issue_ops = [&](int j) {        // operands of the output row of step j: z = z0 + j - 2R
            const int z = z0 + j - 2 * R;
            if (z < z0 || z >= z1 || !active) return;
            T *st = s_op + (j % L::OS) * OPSZ;
            const long long i = (long long)z * g.nx + x;
            if (FWD) {
                f3_cp16(st + lane * VEC, u0 + i);
                f3_cp16(st + L::W + lane * VEC, u1 + i);
                f3_cp16(st + 2 * L::W + lane * VEC, u2 + i);
            } else {
                f3_cp16(st + VEC + lane * VEC, u1 + i);
                if (lane == 0 && x > 0) f3_cp16(st, u1 + i - VEC);
                f3_cp16(st + L::U1LEN + lane * VEC, u2 + i);
                if (!first) f3_cp16(st + L::U1LEN + L::W + lane * VEC, vhat + i);
            }
        };

        const int steps = (z1 - z0) + 2 * R;
        V ring[2 * R + 1];
#pragma unroll
        for (int i = 0; i <= 2 * R; ++i) ring[i] = vec_zero<T, VEC>();
        V u2_prev = vec_zero<T, VEC>();      // adjoint: u2 of row z - 1 (zero above the first row)
        if (!FWD && active && z0 > 0) u2_prev = vec_load<T, VEC>(u2 + (long long)(z0 - 1) * g.nx + x);
#pragma unroll
        for (int j = 0; j < F2_P; ++j) {
            if (j < steps) {
                issue_raw(j);
                issue_ops(j);
            }
            f3_commit();
        }
        for (int j = 0; j < steps; ++j) {
            f3_wait<F2_P - 1>();             // row j (issued P steps ago) has landed for this lane
            __syncwarp();                    // ... and for the whole warp; every lane is done with step j - 1
            if (j + F2_P < steps) {
                issue_raw(j + F2_P);         // overwrites the slot of row j - R - 1: last read in step j - 1
                issue_ops(j + F2_P);
            }
            f3_commit();
            // ---- blur along x straight from the staged row ------------------------------------------------------
            const T *raw = s_raw + (j % L::S) * L::LEN;
            V hx = vec_zero<T, VEC>();
            if (active) {
                T val[VEC + 2 * L::A];
#pragma unroll
                for (int q = 0; q < (VEC + 2 * L::A) / VEC; ++q) {
                    const V w = vec_load<T, VEC>(raw + lane * VEC + q * VEC);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) val[q * VEC + v] = w.v[v];
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    T s = T(0);
#pragma unroll
                    for (int k = 0; k <= 2 * R; ++k) s = f3_fma(tx.t[k], val[L::A + v + R - k], s);
                    hx.v[v] = s;
                }
            }
#pragma unroll
            for (int i = 0; i < 2 * R; ++i) ring[i] = ring[i + 1];
            ring[2 * R] = hx;                // ring[i] = x-blurred row z - R + i
            if (j < 2 * R || !active) continue;

            const int z = z0 + j - 2 * R;
            const long long i0 = (long long)z * g.nx + x;
            const T *op = s_op + (j % L::OS) * OPSZ;
            V av;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                T s = T(0);
#pragma unroll
                for (int k = 0; k <= 2 * R; ++k) s = f3_fma(tz.t[k], ring[2 * R - k].v[v], s);
                av.v[v] = s;
            }
            if (FWD) {
                const T *rz = s_raw + ((j - R) % L::S) * L::LEN + L::A + lane * VEC;        // raw vhat row z
                const T *rz1 = s_raw + ((j - R + 1) % L::S) * L::LEN + L::A + lane * VEC;   // raw vhat row z + 1
                const V vr = vec_load<T, VEC>(rz);
                const T right = (x + VEC < g.nx) ? rz[VEC] : T(0);                           // zero boundary, not the periodic halo
                V vd = vec_zero<T, VEC>();
                if (z + 1 < g.nz) vd = vec_load<T, VEC>(rz1);
                V a0 = vec_load<T, VEC>(op + lane * VEC), a1 = vec_load<T, VEC>(op + L::W + lane * VEC),
                  a2 = vec_load<T, VEC>(op + 2 * L::W + lane * VEC);
                V vc;
#pragma unroll
                for (int v = 0; v < VEC; ++v) vc.v[v] = vr.v[v] * inv_alpha;
                const T rgt = right * inv_alpha;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    T un = (a0.v[v] * inv_beta) * mscale + av.v[v] * inv_alpha;
                    a0.v[v] = un;
                    acc += (double)un * (double)un;
                    const T hi = (v + 1 < VEC) ? vc.v[(v + 1) % VEC] : rgt;
                    const T dx = wx * hi + (-wx) * vc.v[v];
                    un = (a1.v[v] * inv_beta) * mscale + sa * dx;
                    a1.v[v] = un;
                    acc += (double)un * (double)un;
                    const T dz = wz * (vd.v[v] * inv_alpha) + (-wz) * vc.v[v];
                    un = (a2.v[v] * inv_beta) * mscale + sa * dz;
                    a2.v[v] = un;
                    acc += (double)un * (double)un;
                }
                vec_store<T, VEC>(u0 + i0, a0);
                vec_store<T, VEC>(u1 + i0, a1);
                vec_store<T, VEC>(u2 + i0, a2);
            } else {
                const V b1 = vec_load<T, VEC>(op + VEC + lane * VEC);
                const T left = (x > 0) ? op[VEC + lane * VEC - 1] : T(0);
                const V b2 = vec_load<T, VEC>(op + L::U1LEN + lane * VEC);
                V vv = vec_zero<T, VEC>();
                if (!first) vv = vec_load<T, VEC>(op + L::U1LEN + L::W + lane * VEC);
                const T lft = left * inv_beta;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    T r = av.v[v] * inv_beta;
                    const T lo = (v == 0) ? lft : b1.v[(v + VEC - 1) % VEC] * inv_beta;
                    T div = wx * lo + (-wx) * (b1.v[v] * inv_beta);
                    const T loz = (z > 0) ? u2_prev.v[v] * inv_beta : T(0);
                    div = div + (wz * loz + (-wz) * (b2.v[v] * inv_beta));
                    r = r + sa * div;
                    const T vn = first ? r : (vv.v[v] * inv_alpha) * mscale + r;
                    vv.v[v] = vn;
                    acc += (double)vn * (double)vn;
                }
                vec_store<T, VEC>(vhat + i0, vv);
                u2_prev = b2;
            }
        }
        f3_wait<0>();
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[(long long)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

template <typename T, int R, bool FWD>
static int fused2d_v2_launch_r(nsol_lsmr_plan *pl, int first, cudaStream_t s, int *nparts) {
    constexpr int VEC = FastvCfg<T>::VEC;
    using L = F2<T, VEC, R>;
    const GridView &gv = pl->gv;
    Fused2dGeom g;
    g.nx = gv.nx;
    g.nz = gv.nz;
    g.n = gv.n;
    g.zc = fused2d_rows_per_chunk(pl, VEC);
    const dim3 grid((gv.nx + L::W * F2_WARPS - 1) / (L::W * F2_WARPS), (gv.nz + g.zc - 1) / g.zc, 1);
    const size_t smem = L::smem(FWD);
    static bool configured[64] = {false};
    const int dev = pl->ctx->device & 63;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fused2d_v2_kernel<T, R, VEC, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return nsol_fail(pl->ctx, NSOL_ECUDA, "fused2d: smem opt-in %zu -> %s", smem, cudaGetErrorString(e));
        configured[dev] = true;
    }
    // numpy axis 0 = rows (kernel z), axis 1 = columns (x); derivative component 0 acts on x, component 1 on the rows
    fused2d_v2_kernel<T, R, VEC, FWD><<<grid, 32 * F2_WARPS, smem, s>>>(g, (T)gv.w[0], (T)gv.w[1], pl->S, lsq_taps_r<T, R>(pl, 1), lsq_taps_r<T, R>(pl, 0),
                                                                        FWD ? (const T *)pl->v : (const T *)pl->u, (T *)pl->u, (T *)pl->v, pl->part, first);
    *nparts = (int)(grid.x * grid.y);
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

template <typename T>
static int fused2d_v2_launch(nsol_lsmr_plan *pl, bool forward, int first, cudaStream_t s, int *nparts) {
    const int r = pl->desc.radius[0];
#define F2_CASE(RR)                                                                    \
    case RR:                                                                           \
        return forward ? fused2d_v2_launch_r<T, RR, true>(pl, first, s, nparts)        \
                       : fused2d_v2_launch_r<T, RR, false>(pl, first, s, nparts);
    switch (r) {
        F2_CASE(1)
        F2_CASE(2)
        F2_CASE(3)
        F2_CASE(4)
        F2_CASE(5)
        F2_CASE(6)
    }
#undef F2_CASE
    return nsol_fail(pl->ctx, NSOL_EINVAL, "fused2d: radius %d not instantiated", r);
}
