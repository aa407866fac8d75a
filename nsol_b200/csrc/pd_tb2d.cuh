// pd_tb2d.cuh -- 2-D primal-dual iterations with TEMPORAL BLOCKING: K iterations per pass over the state.
//
// A 2-D problem moves 9 words per pixel and iteration through L2 / HBM when every iteration is one pass (pd_iter_kernel), and a
// small single image (BASELINE configs 1, 2) pays one dependent round trip through L2 plus a launch or grid barrier per
// iteration (4.9 us at 256^2, profiles/r2_latency_configs.md).  Here a CTA loads its tile of x, xbar, p_x, p_z, b PLUS a halo of
// K pixels, runs K iterations on chip and stores the tile: 1/K of the passes and launches, at the price of recomputing the halo
// rings (after iteration s of a pass, x / xbar are exact on tile + (K - 1 - s) rings; the rest of the region is scratch).
//
// Everything a pixel owns stays in REGISTERS for the whole pass: the region is 64 columns wide (a lane owns columns lane and
// lane + 32) and nwarps * NR rows high (a warp owns NR consecutive rows), so the x-neighbours come from warp shuffles and the
// z-neighbours from the thread's own registers -- only the first xbar row and the last p'_z row of every warp cross to the
// neighbouring warp through shared memory (two block barriers per iteration).  The first version kept the five arrays in shared
// memory (15 LDS / STS per pixel and iteration) and was bound by exactly that (profiles/r2_tb2d.md).
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
    int halo, tw, th;        // rings around the tw x th tile: tw + 2 halo = 64, th + 2 halo = nwarps * NR
    T wx, wz;
};

template <int NR>
struct PdTbGeom {
    static constexpr int THREADS = NR >= 4 ? 256 : 512;
    static constexpr int ROWS = (THREADS / 32) * NR;       // region height
};

// the K iterations of a pass on the register-resident region.  CHECK = false: the whole region lies inside the image, no pixel
// needs the "outside pixels stay zero" select (most tiles of a large image).
template <typename T, int REG, int DATA, bool UNIT, int NR, bool CHECK, int NW>
__device__ __forceinline__ void pd_tb2d_iterations(const PdTbArgs<T> &a, const int bz, const int lane, const int warp, T (&x)[NR][2],
                                                   T (&bv)[NR][2], T (&px)[NR][2], T (&pz)[NR][2], T (&xb)[NR][2],
                                                   const bool (&in)[NR][2], T (*s_xbf)[64], T (*s_pzl)[64]) {
    const T wx = a.wx, wz = a.wz;
    for (int s = 0; s < a.ksub; ++s) {
        const double *srow = a.sched + ((long long)(a.it + s) * a.batch + bz) * 8;
        const T sigma = (T)srow[0], tau = (T)srow[1], tl = (T)srow[2], theta = (T)srow[3];
        const ConstDiv<T> div_g((T)srow[4]), div_f((T)srow[5]);
        // ---- dual update (primal_dual_solver.py:242-243)
        const T below0 = s_xbf[warp + 1][lane], below1 = s_xbf[warp + 1][lane + 32];
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            // right neighbours: column + 1 (lane 31 of the first half continues in lane 0 of the second; column 64 is outside)
            const T n0 = shfl_down_t(xb[k][0], 1), n1 = shfl_down_t(xb[k][1], 1), f1 = shfl_t(xb[k][1], 0);
            const T hx0 = lane == 31 ? f1 : n0, hx1 = lane == 31 ? T(0) : n1;
            const T hz0 = k + 1 < NR ? xb[(k + 1) % NR][0] : below0, hz1 = k + 1 < NR ? xb[(k + 1) % NR][1] : below1;
            const T qx0 = dual_update<T, REG, UNIT>(px[k][0], hx0, xb[k][0], wx, sigma, div_g);
            const T qx1 = dual_update<T, REG, UNIT>(px[k][1], hx1, xb[k][1], wx, sigma, div_g);
            const T qz0 = dual_update<T, REG, UNIT>(pz[k][0], hz0, xb[k][0], wz, sigma, div_g);
            const T qz1 = dual_update<T, REG, UNIT>(pz[k][1], hz1, xb[k][1], wz, sigma, div_g);
            px[k][0] = (!CHECK || in[k][0]) ? qx0 : px[k][0];
            px[k][1] = (!CHECK || in[k][1]) ? qx1 : px[k][1];
            pz[k][0] = (!CHECK || in[k][0]) ? qz0 : pz[k][0];
            pz[k][1] = (!CHECK || in[k][1]) ? qz1 : pz[k][1];
        }
        s_pzl[warp + 1][lane] = pz[NR - 1][0];
        s_pzl[warp + 1][lane + 32] = pz[NR - 1][1];
        __syncthreads();
        // ---- primal update + over-relaxation (primal_dual_solver.py:246-253)
        const T above0 = s_pzl[warp][lane], above1 = s_pzl[warp][lane + 32];
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            // left neighbours: column - 1 (lane 0 of the second half continues from lane 31 of the first; column -1 is outside)
            const T m0 = shfl_up_t(px[k][0], 1), m1 = shfl_up_t(px[k][1], 1), l0 = shfl_t(px[k][0], 31);
            const T lx0 = lane == 0 ? T(0) : m0, lx1 = lane == 0 ? l0 : m1;
            const T lz0 = k > 0 ? pz[(k + NR - 1) % NR][0] : above0, lz1 = k > 0 ? pz[(k + NR - 1) % NR][1] : above1;
            T d0 = wdiff<T, UNIT>(wx, lx0, px[k][0]);                    // Dx^T p_x
            T d1 = wdiff<T, UNIT>(wx, lx1, px[k][1]);
            d0 = d0 + wdiff<T, UNIT>(wz, lz0, pz[k][0]);                 // += Dz^T p_z
            d1 = d1 + wdiff<T, UNIT>(wz, lz1, pz[k][1]);
            T xn0, xbn0, xn1, xbn1;
            primal_update<T, DATA>(x[k][0], bv[k][0], d0, tau, tl, theta, div_f, xn0, xbn0);
            primal_update<T, DATA>(x[k][1], bv[k][1], d1, tau, tl, theta, div_f, xn1, xbn1);
            x[k][0] = (!CHECK || in[k][0]) ? xn0 : x[k][0];
            x[k][1] = (!CHECK || in[k][1]) ? xn1 : x[k][1];
            xb[k][0] = (!CHECK || in[k][0]) ? xbn0 : xb[k][0];
            xb[k][1] = (!CHECK || in[k][1]) ? xbn1 : xb[k][1];
        }
        s_xbf[warp][lane] = xb[0][0];
        s_xbf[warp][lane + 32] = xb[0][1];
        __syncthreads();
    }

}

template <typename T, int REG, int DATA, bool UNIT, int NR>
__global__ void __launch_bounds__(PdTbGeom<NR>::THREADS, 2) pd_tb2d_kernel(const PdTbArgs<T> a) {
    constexpr int NW = PdTbGeom<NR>::THREADS / 32;
    __shared__ T s_xbf[NW + 1][64];      // first xbar row of every warp (+ a zero row below the region)
    __shared__ T s_pzl[NW + 1][64];      // last p'_z row of every warp, stored at [warp + 1] (+ a zero row above the region)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = a.halo;
    const int c0 = (int)blockIdx.x * a.tw - H + lane;                 // image column of this lane's first pixel; second: + 32
    const int r0 = (int)blockIdx.y * a.th - H + warp * NR;            // image row of this warp's first row
    const int bz = blockIdx.z;
    const long long base = (long long)bz * a.n, bbase = (long long)bz * a.b_stride;

    T x[NR][2], bv[NR][2], px[NR][2], pz[NR][2], xb[NR][2];
    bool in[NR][2];
#pragma unroll
    for (int k = 0; k < NR; ++k)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = r0 + k, c = c0 + 32 * h;
            in[k][h] = r >= 0 && r < a.nz && c >= 0 && c < a.nx;
            const long long o = (long long)r * a.nx + c;
            xb[k][h] = in[k][h] ? a.xbar_in[base + o] : T(0);
            x[k][h] = in[k][h] ? a.x_in[base + o] : T(0);
            bv[k][h] = in[k][h] ? a.b[bbase + o] : T(0);
            px[k][h] = in[k][h] ? a.px_in[base + o] : T(0);
            pz[k][h] = in[k][h] ? a.pz_in[base + o] : T(0);
        }
    if (warp == 0) {
        s_xbf[NW][lane] = s_xbf[NW][lane + 32] = T(0);
        s_pzl[0][lane] = s_pzl[0][lane + 32] = T(0);
    }
    s_xbf[warp][lane] = xb[0][0];
    s_xbf[warp][lane + 32] = xb[0][1];
    __syncthreads();

    const bool inside = r0 - warp * NR >= 0 && r0 - warp * NR + NW * NR <= a.nz && c0 - lane >= 0 && c0 - lane + 64 <= a.nx;   // CTA-uniform
    if (inside) pd_tb2d_iterations<T, REG, DATA, UNIT, NR, false, NW>(a, bz, lane, warp, x, bv, px, pz, xb, in, s_xbf, s_pzl);
    else pd_tb2d_iterations<T, REG, DATA, UNIT, NR, true, NW>(a, bz, lane, warp, x, bv, px, pz, xb, in, s_xbf, s_pzl);

    // ---- store the pixels of the tile
#pragma unroll
    for (int k = 0; k < NR; ++k)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int rr = warp * NR + k, cc = lane + 32 * h;            // region coordinates
            if (in[k][h] && rr >= H && rr < H + a.th && cc >= H && cc < H + a.tw) {
                const long long o = base + (long long)(r0 + k) * a.nx + c0 + 32 * h;
                a.xbar_out[o] = xb[k][h];
                a.x_out[o] = x[k][h];
                a.px_out[o] = px[k][h];
                a.pz_out[o] = pz[k][h];
            }
        }
}

template <typename T, int R, int D, bool UNIT>
static int pd_tb2d_launch_one(nsol_ctx *ctx, int nr, const PdTbArgs<T> &a, dim3 grid, cudaStream_t s) {
    (void)ctx;
    if (nr == 1) pd_tb2d_kernel<T, R, D, UNIT, 1><<<grid, PdTbGeom<1>::THREADS, 0, s>>>(a);
    else if (nr == 2) pd_tb2d_kernel<T, R, D, UNIT, 2><<<grid, PdTbGeom<2>::THREADS, 0, s>>>(a);
    else if (nr == 4) pd_tb2d_kernel<T, R, D, UNIT, 4><<<grid, PdTbGeom<4>::THREADS, 0, s>>>(a);
    else return NSOL_EINVAL;
    return NSOL_OK;
}

// region height of the variant with nr rows per thread (1, 2 or 4); 0: no such variant
static inline int pd_tb2d_rows(int nr) { return nr == 1 ? PdTbGeom<1>::ROWS : nr == 2 ? PdTbGeom<2>::ROWS : nr == 4 ? PdTbGeom<4>::ROWS : 0; }

template <typename T>
static int pd_tb2d_launch(nsol_ctx *ctx, int reg, int data, int nr, const PdTbArgs<T> &a, dim3 grid, cudaStream_t s) {
    const bool unit = a.wx == T(1) && a.wz == T(1);
#define NSOL_PD_CASE(R, D)                                                                  \
    if (reg == R && data == D)                                                              \
        return unit ? pd_tb2d_launch_one<T, R, D, true>(ctx, nr, a, grid, s)                \
                    : pd_tb2d_launch_one<T, R, D, false>(ctx, nr, a, grid, s);
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L1)
#undef NSOL_PD_CASE
    return NSOL_EINVAL;
}
