// pd_tb2d.cuh -- 2-D primal-dual iterations with TEMPORAL BLOCKING: K iterations per pass over the state.
//
// A 2-D problem moves 9 words per pixel and iteration through L2 / HBM when every iteration is one pass (pd_iter_kernel), and a
// small single image (BASELINE configs 1, 2) pays one dependent round trip through L2 plus a launch or grid barrier per
// iteration (4.9 us at 256^2, profiles/r2_latency_configs.md).  Here a CTA loads its tile of x, xbar, p_x, p_z, b PLUS a halo of
// K pixels into shared memory, runs K iterations there (two block barriers each: dual update in place, then primal update in
// place) on a region that shrinks by one ring per iteration, and stores the tile: 1/K of the passes, 1/K of the launches, at the
// price of recomputing the halo rings.  Iteration s of a launch computes p' on tile + (K - s) rings and x, xbar on tile +
// (K - 1 - s) rings; whatever lies outside those boxes is stale and never read by a box that matters.
//
// Same per-pixel arithmetic as pd_iter_body (dual_update / primal_update, the same operand order in the divergence), and the same
// boundary rules: pixels outside the image hold xbar = 0 and p = 0 and are never updated -- the reference's mode="constant"
// differences (nsol/linear_operators.py:98-106).  The float64 path therefore stays bit-identical to numpy.
// State arrays ping-pong between launches (neighbouring tiles read each other's halo pixels of the previous pass), which for x
// needs the plan's second x array (nsol_pd_plan::x_alt).
#pragma once

template <typename T>
struct PdTbArgs {
    const T *xbar_in, *x_in, *px_in, *pz_in, *b;
    T *xbar_out, *x_out, *px_out, *pz_out;
    const double *sched;     // [it][batch][8]
    long long n, b_stride;
    int nx, nz, batch;
    int it, ksub;            // first iteration of this launch, iterations in this launch (<= halo)
    int halo, tw, th;        // rings loaded around the tw x th tile
    T wx, wz;
};

template <typename T, int REG, int DATA, bool UNIT>
__global__ void __launch_bounds__(256, 2) pd_tb2d_kernel(const PdTbArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = a.halo, RW = a.tw + 2 * H, RH = a.th + 2 * H;
    T *s_xb = reinterpret_cast<T *>(smem_raw);
    T *s_x = s_xb + RW * RH;
    T *s_b = s_x + RW * RH;
    T *s_px = s_b + RW * RH;
    T *s_pz = s_px + RW * RH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int c_org = (int)blockIdx.x * a.tw - H, r_org = (int)blockIdx.y * a.th - H;
    const int bz = blockIdx.z;
    const long long base = (long long)bz * a.n, bbase = (long long)bz * a.b_stride;

    // ---- load tile + halo; outside the image everything is zero and stays zero
    for (int rr = warp; rr < RH; rr += nw) {
        const int r = r_org + rr;
        const bool row_in = r >= 0 && r < a.nz;
        const long long ro = (long long)r * a.nx;
        for (int cc = lane; cc < RW; cc += 32) {
            const int c = c_org + cc;
            const bool in = row_in && c >= 0 && c < a.nx;
            const int i = rr * RW + cc;
            const long long o = ro + c;
            s_xb[i] = in ? a.xbar_in[base + o] : T(0);
            s_x[i] = in ? a.x_in[base + o] : T(0);
            s_b[i] = in ? a.b[bbase + o] : T(0);
            s_px[i] = in ? a.px_in[base + o] : T(0);
            s_pz[i] = in ? a.pz_in[base + o] : T(0);
        }
    }
    __syncthreads();

    // image bounds in region coordinates
    const int rr_lo = max(0, -r_org), rr_hi = min(RH, a.nz - r_org);
    const int cc_lo = max(0, -c_org), cc_hi = min(RW, a.nx - c_org);
    const T wx = a.wx, wz = a.wz;
    for (int s = 0; s < a.ksub; ++s) {
        const double *srow = a.sched + ((long long)(a.it + s) * a.batch + bz) * 8;
        const T sigma = (T)srow[0], tau = (T)srow[1], tl = (T)srow[2], theta = (T)srow[3];
        const ConstDiv<T> div_g((T)srow[4]), div_f((T)srow[5]);
        // ---- dual update on tile + m rings (primal_dual_solver.py:242-243)
        const int m = a.ksub - s;
        {
            const int r0 = max(rr_lo, H - m), r1 = min(rr_hi, H + a.th + m);
            const int c0 = max(cc_lo, H - m), c1 = min(cc_hi, H + a.tw + m);
            for (int rr = r0 + warp; rr < r1; rr += nw)
                for (int cc = c0 + lane; cc < c1; cc += 32) {
                    const int i = rr * RW + cc;
                    const T xbc = s_xb[i];
                    const T hx = cc + 1 < RW ? s_xb[i + 1] : T(0);
                    const T hz = rr + 1 < RH ? s_xb[i + RW] : T(0);
                    s_px[i] = dual_update<T, REG, UNIT>(s_px[i], hx, xbc, wx, sigma, div_g);
                    s_pz[i] = dual_update<T, REG, UNIT>(s_pz[i], hz, xbc, wz, sigma, div_g);
                }
        }
        __syncthreads();
        // ---- primal update + over-relaxation on tile + (m - 1) rings (primal_dual_solver.py:246-253)
        {
            const int r0 = max(rr_lo, H - m + 1), r1 = min(rr_hi, H + a.th + m - 1);
            const int c0 = max(cc_lo, H - m + 1), c1 = min(cc_hi, H + a.tw + m - 1);
            for (int rr = r0 + warp; rr < r1; rr += nw)
                for (int cc = c0 + lane; cc < c1; cc += 32) {
                    const int i = rr * RW + cc;
                    const T pxc = s_px[i], pzc = s_pz[i];
                    const T lx = cc > 0 ? s_px[i - 1] : T(0);
                    const T lz = rr > 0 ? s_pz[i - RW] : T(0);
                    T div = wdiff<T, UNIT>(wx, lx, pxc);                 // Dx^T p_x
                    div = div + wdiff<T, UNIT>(wz, lz, pzc);             // += Dz^T p_z
                    T xn, xbn;
                    primal_update<T, DATA>(s_x[i], s_b[i], div, tau, tl, theta, div_f, xn, xbn);
                    s_x[i] = xn;
                    s_xb[i] = xbn;
                }
        }
        __syncthreads();
    }

    // ---- store the tile
    for (int rr = H + warp; rr < H + a.th; rr += nw) {
        const int r = r_org + rr;
        if (r >= a.nz) break;
        const long long ro = base + (long long)r * a.nx;
        for (int cc = H + lane; cc < H + a.tw; cc += 32) {
            const int c = c_org + cc;
            if (c >= a.nx) break;
            const int i = rr * RW + cc;
            a.xbar_out[ro + c] = s_xb[i];
            a.x_out[ro + c] = s_x[i];
            a.px_out[ro + c] = s_px[i];
            a.pz_out[ro + c] = s_pz[i];
        }
    }
}

template <typename T, int R, int D, bool UNIT>
static int pd_tb2d_launch_one(nsol_ctx *ctx, const PdTbArgs<T> &a, dim3 grid, size_t smem, cudaStream_t s) {
    static size_t configured[64] = {0};
    const int dev = ctx->device & 63;
    if (smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pd_tb2d_kernel<T, R, D, UNIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return nsol_fail(ctx, NSOL_ECUDA, "pd tb2d: smem opt-in %zu -> %s", smem, cudaGetErrorString(e));
        configured[dev] = smem;
    }
    pd_tb2d_kernel<T, R, D, UNIT><<<grid, 256, smem, s>>>(a);
    return NSOL_OK;
}

template <typename T>
static int pd_tb2d_launch(nsol_ctx *ctx, int reg, int data, const PdTbArgs<T> &a, dim3 grid, size_t smem, cudaStream_t s) {
    const bool unit = a.wx == T(1) && a.wz == T(1);
#define NSOL_PD_CASE(R, D)                                                                  \
    if (reg == R && data == D)                                                              \
        return unit ? pd_tb2d_launch_one<T, R, D, true>(ctx, a, grid, smem, s)              \
                    : pd_tb2d_launch_one<T, R, D, false>(ctx, a, grid, smem, s);
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L1)
#undef NSOL_PD_CASE
    return NSOL_EINVAL;
}
