// lsmr_fastv.cuh -- vectorised, radius-specialised versions of the row-mapped LSMR vector kernels
// (fast_blur_pass_kernel / fast_fwd_kernel / fast_adj_kernel / lsmr_update_kernel in lsmr_kernels.cu).
//
// ncu on the generic kernels at 4096^2 float64 (profiles/r1_lsmr_v2.md): the blur pass was issue-bound
// (224 instructions per thread and element: run-time tap loop, 64-bit index products, taps fetched from the
// parameter bank by a dynamic index), fwd / adj spent 390 instructions per element with the L1 path 70 % busy,
// and the update kernel kept only ~4 scalar loads per thread in flight (52 % of DRAM).  Here:
//   * the radius is a template parameter (taps unrolled, static operands), index math is 32-bit inside a row,
//   * every thread owns VEC = 16 bytes of consecutive x (128-bit loads / stores), the x-blur of its VEC outputs
//     reads VEC + 2R inputs once instead of VEC (2R + 1),
//   * the update kernel streams two 128-bit vectors per array and thread.
// Per element the arithmetic is that of the generic kernels (same tap order, same operation order); only the
// partial sums of the norms are grouped differently.  The generic kernels remain the fallback for radii >
// FASTV_MAX_R, odd nx and axes shorter than the mask.
#pragma once

#define FASTV_MAX_R 6

template <typename T, int R>
struct TapsR {
    T t[2 * R + 1];
};

template <typename T, int R>
static TapsR<T, R> lsq_taps_r(const nsol_lsmr_plan *pl, int ax) {
    TapsR<T, R> t;
    for (int k = 0; k <= 2 * R; ++k) t.t[k] = (R > 0 && ax >= 0) ? (T)pl->taps[ax][k] : T(1);
    return t;
}

// geometry the vector kernels need (32-bit inside a plane)
template <typename T>
struct FastvGeom {
    int nx, ny, nz, dim;
    long long n;
    T wx, wy, wz;
    int b_op;
    int slab, grad_lo, grad_hi, ghost;
};

template <typename T>
static FastvGeom<T> make_fastv_geom(const LsqGeom<T> &g) {
    FastvGeom<T> f;
    f.nx = g.nx;
    f.ny = g.ny;
    f.nz = g.nz;
    f.dim = g.dim;
    f.n = g.n;
    f.wx = g.w[0];
    f.wy = g.dim == 3 ? g.w[1] : T(0);
    f.wz = g.dim >= 2 ? g.w[g.dim - 1] : T(0);
    f.b_op = g.b_op;
    f.slab = g.slab;
    f.grad_lo = g.grad_lo;
    f.grad_hi = g.grad_hi;
    f.ghost = g.ghost;
    return f;
}

// ---- separable pass along y (kaxis 1) or z (kaxis 2), periodic; z-slab: halo planes instead of the wrap ----
// Every thread produces FASTV_ROWS consecutive outputs of its line from one sweep over FASTV_ROWS + 2R inputs
// (all loads issued up front): (ROWS + 2R) / ROWS loads per output instead of 2R + 1.
#define FASTV_ROWS 8
// (bx, by, bz): block index of the row-mapped grid -- the kernel's blockIdx, or a virtual index in the persistent
// cooperative solve (lsmr_coopv_kernel)
template <typename T, int R, int VEC>
__device__ __forceinline__ void fastv_blur_pass_body(const FastvGeom<T> &g, const TapsR<T, R> &tp, int kaxis, const T *__restrict__ in,
                                                     T *__restrict__ out, const T *__restrict__ halo_lo, const T *__restrict__ halo_hi,
                                                     unsigned bx, unsigned by, unsigned bz) {
    using V = Vec<T, VEC>;
    const int x = (int)(bx * FAST_TH + threadIdx.x) * VEC;
    if (x >= g.nx) return;
    // the grid axis of the blurred direction counts groups of FASTV_ROWS positions
    const int y = kaxis == 1 ? 0 : (int)by, z = kaxis == 2 ? 0 : (int)bz;
    const int pos0 = (kaxis == 1 ? (int)by : (int)bz) * FASTV_ROWS;
    const long long plane = (long long)g.nx * g.ny;
    const long long st = kaxis == 1 ? (long long)g.nx : plane;
    const int ext = kaxis == 1 ? g.ny : g.nz;
    const long long line = (long long)z * plane + (long long)y * g.nx + x;    // element at position 0 of this line
    const T *base = in + line;
    const long long po = (long long)y * g.nx + x;            // offset inside a plane (z-slab halos)
    const bool halo = kaxis == 2 && g.slab;
    V val[FASTV_ROWS + 2 * R];
#pragma unroll
    for (int j = 0; j < FASTV_ROWS + 2 * R; ++j) {
        int q = pos0 - R + j;                                // input position (before the wrap)
        const T *src;
        if (halo) {
            src = q < 0 ? halo_lo + (long long)(q + g.ghost) * st + po
                        : (q >= g.nz ? halo_hi + (long long)(q - g.nz < g.ghost ? q - g.nz : g.ghost - 1) * st + po : base + (long long)q * st);
        } else {
            q += q < 0 ? ext : 0;
            if (q >= ext) {
                q -= ext;
                if (q >= ext) q %= ext;                      // only inputs of a partial group's unused outputs get here
            }
            src = base + (long long)q * st;
        }
        val[j] = vec_load<T, VEC>(src);
    }
#pragma unroll
    for (int o = 0; o < FASTV_ROWS; ++o) {
        if (pos0 + o < ext) {
            V acc = vec_zero<T, VEC>();
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {               // input position (pos0 + o) - (k - R): same tap order as fast_blur_line
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc.v[v] += tp.t[k] * val[o + 2 * R - k].v[v];
            }
            vec_store<T, VEC>(out + line + (long long)(pos0 + o) * st, acc);
        }
    }
}

template <typename T, int R, int VEC>
__global__ void __launch_bounds__(FAST_TH) fastv_blur_pass_kernel(FastvGeom<T> g, TapsR<T, R> tp, int kaxis, const T *__restrict__ in,
                                                                  T *__restrict__ out, const T *__restrict__ halo_lo,
                                                                  const T *__restrict__ halo_hi) {
    fastv_blur_pass_body<T, R, VEC>(g, tp, kaxis, in, out, halo_lo, halo_hi, blockIdx.x, blockIdx.y, blockIdx.z);
}

// x-blur of VEC consecutive outputs of one row (periodic): reads the aligned 128-bit vectors covering
// [x - R, x + VEC - 1 + R]; tap order as fast_blur_line
template <typename T, int R, int VEC>
__device__ __forceinline__ Vec<T, VEC> fastv_blur_x(const TapsR<T, R> &tp, const T *__restrict__ row, int x, int nx) {
    Vec<T, VEC> out;
    if (R == 0) return vec_load<T, VEC>(row + x);
    constexpr int A = (R + VEC - 1) / VEC * VEC;             // halo rounded up to whole vectors
    constexpr int NV = (VEC + 2 * A) / VEC;
    T val[VEC + 2 * A];
    if (x - A >= 0 && x + VEC + A <= nx) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const Vec<T, VEC> w = vec_load<T, VEC>(row + x - A + j * VEC);
#pragma unroll
            for (int v = 0; v < VEC; ++v) val[j * VEC + v] = w.v[v];
        }
    } else {
#pragma unroll
        for (int j = A - R; j < VEC + A + R; ++j) {
            int q = x - A + j;
            q += q < 0 ? nx : 0;
            q -= q >= nx ? nx : 0;
            val[j] = row[q];
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k <= 2 * R; ++k) acc += tp.t[k] * val[A + v + R - k];     // input position (x + v) - (k - R)
        out.v[v] = acc;
    }
    return out;
}

// u <- (u * inv_beta) * (-alpha) + [A v; sqrt_alpha B v], v = vhat * inv_alpha, partial ||u||^2   (lsmr.py:336-338)
template <typename T, int R, int VEC>
__device__ __forceinline__ void fastv_fwd_body(const FastvGeom<T> &g, const LsmrScalars *S, const TapsR<T, R> &tx, const T *__restrict__ src,
                                               const T *__restrict__ vhat, T *__restrict__ u, const T *__restrict__ v_hi, unsigned bx,
                                               unsigned by, unsigned bz, double &acc) {
    using V = Vec<T, VEC>;
    const int x = (int)(bx * FAST_TH + threadIdx.x) * VEC;
    const int y = (int)by, z = (int)bz;
    if (x < g.nx) {
        const T inv_alpha = (T)S->inv_alpha, inv_beta = (T)S->inv_beta, malpha = (T)(-S->alpha), sa = (T)S->sqrt_alpha;
        const long long plane = (long long)g.nx * g.ny;
        const long long rowoff = (long long)z * plane + (long long)y * g.nx;
        const long long i = rowoff + x;
        const V hx = fastv_blur_x<T, R, VEC>(tx, src + rowoff, x, g.nx);
        V u0 = vec_load<T, VEC>(u + i);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const T un = (u0.v[v] * inv_beta) * malpha + hx.v[v] * inv_alpha;
            u0.v[v] = un;
            acc += (double)un * (double)un;
        }
        vec_store<T, VEC>(u + i, u0);
        if (g.b_op == NSOL_B_GRAD) {
            const V vr = vec_load<T, VEC>(vhat + i);
            V vc;
#pragma unroll
            for (int v = 0; v < VEC; ++v) vc.v[v] = vr.v[v] * inv_alpha;
            // component 0: along x
            {
                const T right = (x + VEC < g.nx) ? vhat[i + VEC] * inv_alpha : T(0);
                T *uk = u + g.n;
                V uv = vec_load<T, VEC>(uk + i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const T hi = (v + 1 < VEC) ? vc.v[(v + 1) % VEC] : right;
                    const T dk = g.wx * hi + (-g.wx) * vc.v[v];
                    const T un = (uv.v[v] * inv_beta) * malpha + sa * dk;
                    uv.v[v] = un;
                    acc += (double)un * (double)un;
                }
                vec_store<T, VEC>(uk + i, uv);
            }
            if (g.dim == 3) {   // component 1: along y
                V hv = vec_zero<T, VEC>();
                if (y + 1 < g.ny) hv = vec_load<T, VEC>(vhat + i + g.nx);
                T *uk = u + 2 * g.n;
                V uv = vec_load<T, VEC>(uk + i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const T hi = (y + 1 < g.ny) ? hv.v[v] * inv_alpha : T(0);
                    const T dk = g.wy * hi + (-g.wy) * vc.v[v];
                    const T un = (uv.v[v] * inv_beta) * malpha + sa * dk;
                    uv.v[v] = un;
                    acc += (double)un * (double)un;
                }
                vec_store<T, VEC>(uk + i, uv);
            }
            if (g.dim >= 2) {   // last component: along z (the slowest axis)
                V hv = vec_zero<T, VEC>();
                bool have = z + 1 < g.nz;
                if (have) hv = vec_load<T, VEC>(vhat + i + plane);
                else if (g.slab && g.grad_hi) {
                    hv = vec_load<T, VEC>(v_hi + (long long)y * g.nx + x);
                    have = true;
                }
                T *uk = u + (long long)g.dim * g.n;
                V uv = vec_load<T, VEC>(uk + i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const T hi = have ? hv.v[v] * inv_alpha : T(0);
                    const T dk = g.wz * hi + (-g.wz) * vc.v[v];
                    const T un = (uv.v[v] * inv_beta) * malpha + sa * dk;
                    uv.v[v] = un;
                    acc += (double)un * (double)un;
                }
                vec_store<T, VEC>(uk + i, uv);
            }
        } else if (g.b_op == NSOL_B_IDENTITY) {
            const V vr = vec_load<T, VEC>(vhat + i);
            T *uk = u + g.n;
            V uv = vec_load<T, VEC>(uk + i);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const T un = (uv.v[v] * inv_beta) * malpha + sa * (vr.v[v] * inv_alpha);
                uv.v[v] = un;
                acc += (double)un * (double)un;
            }
            vec_store<T, VEC>(uk + i, uv);
        }
    }
}

template <typename T, int R, int VEC>
__global__ void __launch_bounds__(FAST_TH) fastv_fwd_kernel(FastvGeom<T> g, const LsmrScalars *__restrict__ S, TapsR<T, R> tx,
                                                            const T *__restrict__ src, const T *__restrict__ vhat, T *__restrict__ u,
                                                            double *__restrict__ part, const T *__restrict__ v_hi) {
    if (S->done) return;
    double acc = 0.0;
    fastv_fwd_body<T, R, VEC>(g, S, tx, src, vhat, u, v_hi, blockIdx.x, blockIdx.y, blockIdx.z, acc);
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
}

// vhat <- (vhat * inv_alpha) * (-beta) + (A^T u0 + sqrt_alpha B^T u1..), u = uhat * inv_beta, partial ||v||^2 (lsmr.py:342-344)
template <typename T, int R, int VEC>
__device__ __forceinline__ void fastv_adj_body(const FastvGeom<T> &g, const LsmrScalars *S, const TapsR<T, R> &tx, const T *__restrict__ src,
                                               const T *__restrict__ u, T *__restrict__ vhat, int first, const T *__restrict__ uz_lo,
                                               unsigned bx, unsigned by, unsigned bz, double &acc) {
    using V = Vec<T, VEC>;
    const int x = (int)(bx * FAST_TH + threadIdx.x) * VEC;
    const int y = (int)by, z = (int)bz;
    if (x < g.nx) {
        const T inv_alpha = (T)S->inv_alpha, inv_beta = (T)S->inv_beta, mbeta = (T)(-S->beta), sa = (T)S->sqrt_alpha;
        const long long plane = (long long)g.nx * g.ny;
        const long long rowoff = (long long)z * plane + (long long)y * g.nx;
        const long long i = rowoff + x;
        const V hx = fastv_blur_x<T, R, VEC>(tx, src + rowoff, x, g.nx);
        V r;
#pragma unroll
        for (int v = 0; v < VEC; ++v) r.v[v] = hx.v[v] * inv_beta;
        if (g.b_op == NSOL_B_GRAD) {
            V div;
            {   // component 0: along x
                const T *uk = u + g.n;
                const V uv = vec_load<T, VEC>(uk + i);
                const T left = (x > 0) ? uk[i - 1] * inv_beta : T(0);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const T lo = (v == 0) ? left : uv.v[(v + VEC - 1) % VEC] * inv_beta;
                    div.v[v] = g.wx * lo + (-g.wx) * (uv.v[v] * inv_beta);
                }
            }
            if (g.dim == 3) {   // component 1: along y
                const T *uk = u + 2 * g.n;
                const V uv = vec_load<T, VEC>(uk + i);
                V lv = vec_zero<T, VEC>();
                if (y > 0) lv = vec_load<T, VEC>(uk + i - g.nx);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const T lo = (y > 0) ? lv.v[v] * inv_beta : T(0);
                    div.v[v] = div.v[v] + (g.wy * lo + (-g.wy) * (uv.v[v] * inv_beta));
                }
            }
            if (g.dim >= 2) {   // last component: along z
                const T *uk = u + (long long)g.dim * g.n;
                const V uv = vec_load<T, VEC>(uk + i);
                V lv = vec_zero<T, VEC>();
                bool have = z > 0;
                if (have) lv = vec_load<T, VEC>(uk + i - plane);
                else if (g.slab && g.grad_lo) {
                    lv = vec_load<T, VEC>(uz_lo + (long long)y * g.nx + x);
                    have = true;
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const T lo = have ? lv.v[v] * inv_beta : T(0);
                    div.v[v] = div.v[v] + (g.wz * lo + (-g.wz) * (uv.v[v] * inv_beta));
                }
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) r.v[v] = r.v[v] + sa * div.v[v];
        } else if (g.b_op == NSOL_B_IDENTITY) {
            const V uv = vec_load<T, VEC>(u + g.n + i);
#pragma unroll
            for (int v = 0; v < VEC; ++v) r.v[v] = r.v[v] + sa * (uv.v[v] * inv_beta);
        }
        V vv = vec_zero<T, VEC>();
        if (!first) vv = vec_load<T, VEC>(vhat + i);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const T vn = first ? r.v[v] : (vv.v[v] * inv_alpha) * mbeta + r.v[v];
            vv.v[v] = vn;
            acc += (double)vn * (double)vn;
        }
        vec_store<T, VEC>(vhat + i, vv);
    }
}

template <typename T, int R, int VEC>
__global__ void __launch_bounds__(FAST_TH) fastv_adj_kernel(FastvGeom<T> g, const LsmrScalars *__restrict__ S, TapsR<T, R> tx,
                                                            const T *__restrict__ src, const T *__restrict__ u, T *__restrict__ vhat,
                                                            double *__restrict__ part, int first, const T *__restrict__ uz_lo) {
    if (S->done) return;
    double acc = 0.0;
    fastv_adj_body<T, R, VEC>(g, S, tx, src, u, vhat, first, uz_lo, blockIdx.x, blockIdx.y, blockIdx.z, acc);
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
}

// hbar = c_hbar*hbar + h;  x += c_x*hbar;  h = c_h*h + v;  partial ||x||^2   (lsmr.py:373-377, 421)
// two 128-bit vectors per array and thread in flight; n must be a multiple of VEC
template <typename T, int VEC>
__global__ void __launch_bounds__(LSMR_THREADS) fastv_update_kernel(long long nvec, const LsmrScalars *__restrict__ S, const T *__restrict__ vhat,
                                                                    T *__restrict__ h, T *__restrict__ hbar, T *__restrict__ x,
                                                                    double *__restrict__ part) {
    using V = Vec<T, VEC>;
    if (S->done) return;
    const T c_hbar = (T)S->c_hbar, c_x = (T)S->c_x, c_h = (T)S->c_h, inv_alpha = (T)S->inv_alpha;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    auto one = [&](long long j, const V &hv, const V &hb, const V &xv, const V &vv) {
        V hbn, xn, hn;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            hbn.v[v] = hb.v[v] * c_hbar + hv.v[v];
            xn.v[v] = xv.v[v] + c_x * hbn.v[v];
            hn.v[v] = hv.v[v] * c_h + vv.v[v] * inv_alpha;
            acc += (double)xn.v[v] * (double)xn.v[v];
        }
        vec_store<T, VEC>(hbar + j * VEC, hbn);
        vec_store<T, VEC>(x + j * VEC, xn);
        vec_store<T, VEC>(h + j * VEC, hn);
    };
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; j + stride < nvec; j += 2 * stride) {
        const long long j2 = j + stride;
        const V h0 = vec_load<T, VEC>(h + j * VEC), b0 = vec_load<T, VEC>(hbar + j * VEC), x0 = vec_load<T, VEC>(x + j * VEC),
                v0 = vec_load<T, VEC>(vhat + j * VEC);
        const V h1 = vec_load<T, VEC>(h + j2 * VEC), b1 = vec_load<T, VEC>(hbar + j2 * VEC), x1 = vec_load<T, VEC>(x + j2 * VEC),
                v1 = vec_load<T, VEC>(vhat + j2 * VEC);
        one(j, h0, b0, x0, v0);
        one(j2, h1, b1, x1, v1);
    }
    if (j < nvec) {
        const V h0 = vec_load<T, VEC>(h + j * VEC), b0 = vec_load<T, VEC>(hbar + j * VEC), x0 = vec_load<T, VEC>(x + j * VEC),
                v0 = vec_load<T, VEC>(vhat + j * VEC);
        one(j, h0, b0, x0, v0);
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// ---- host-side dispatch -------------------------------------------------------------------------
template <typename T>
struct FastvCfg {
    static constexpr int VEC = 16 / (int)sizeof(T);
};

// radius of the blur along numpy axis ax of the plan (-1: no blur)
static inline int lsq_radius(const nsol_lsmr_plan *pl, int ax) {
    return (pl->desc.a_op == NSOL_A_BLUR && ax >= 0) ? pl->desc.radius[ax] : -1;
}

// the vector kernels apply when rows are a whole number of 16-byte vectors and every mask is shorter than its axis
static bool fastv_ok(const nsol_lsmr_plan *pl) {
    const GridView &gv = pl->gv;
    const int vec = gv.dtype == NSOL_F32 ? 4 : 2;
    if (gv.nx % vec) return false;
    if (pl->ctx->lsmr_path == 3) return false;      // tuning: force the generic kernels
    for (int ax = 0; ax < gv.dim; ++ax) {
        const int r = lsq_radius(pl, ax);
        if (r > FASTV_MAX_R) return false;
        const int ext = (int)pl->grid.shape[ax];
        if (r >= 0 && ext <= r) return false;
    }
    return true;
}

#define FASTV_SWITCH_R(r, CALL)              \
    switch (r) {                             \
    case 0: { constexpr int R = 0; CALL; } break; \
    case 1: { constexpr int R = 1; CALL; } break; \
    case 2: { constexpr int R = 2; CALL; } break; \
    case 3: { constexpr int R = 3; CALL; } break; \
    case 4: { constexpr int R = 4; CALL; } break; \
    case 5: { constexpr int R = 5; CALL; } break; \
    default: { constexpr int R = 6; CALL; } break; \
    }

// =====================================================================================================
// Fused 2-D forward / adjoint kernels: both blur passes inside the consumer (no separate pass kernel, no
// scratch array).  A warp owns a strip of W = 32 * VEC columns and marches down a chunk of rows: every new
// input row is blurred along x through a warp-private shared-memory row (with its periodic halo columns) and
// pushed into a register ring of 2R+1 x-blurred rows; the blur along the rows is a dot product over the ring.
// DRAM traffic is the algorithmic 7 words (forward) / 5 words (adjoint) per pixel; the rows needed a second time
// (v[z], v[z+1] for the gradient; the 2R rows of ring warm-up per chunk) come from L1 / L2.
// The blur is evaluated x-first (the separate-pass path goes rows-first): same taps, different rounding order.
// =====================================================================================================
#define FUSED2D_WARPS 4
#ifndef FUSED2D_MINB
#define FUSED2D_MINB 5   // 96 registers -> 5 CTAs = 20 warps per SM (measured: 4 CTAs 625 us, 5 CTAs 572 us, 6 CTAs with spills 664 us at 4096^2 float64)
#endif

template <typename T, int VEC, int RX>
struct Fused2dRow {
    static constexpr int A = RX == 0 ? 0 : (RX + VEC - 1) / VEC * VEC;   // halo columns, whole vectors
    static constexpr int W = 32 * VEC;
    static constexpr int LEN = W + 2 * A;
};

// Per-lane addressing of one input row of the warp's strip: own VEC columns + (edge lanes) one halo vector.
// The loads of row j+1 are issued before row j is consumed (software pipelining), so the x-blur never waits
// on a global load it has just issued.
template <typename T, int VEC, int RX>
struct Fused2dLane {
    using V = Vec<T, VEC>;
    using L = Fused2dRow<T, VEC, RX>;
    static constexpr int HV = L::A / VEC;                // halo vectors per side
    int x, lane;
    bool active, has_halo;
    int halo_col, halo_slot;                             // column of this lane's halo vector / its slot in the shared row

    __device__ __forceinline__ void init(int xw0, int nx, int lane_) {
        lane = lane_;
        x = xw0 + lane * VEC;
        active = x < nx;
        has_halo = RX > 0 && lane < 2 * HV;
        const int wcols = min(L::W, nx - xw0);           // columns of this warp inside the row
        // lanes 0..HV-1: left halo vectors, HV..2HV-1: right halo vectors (periodic; whole vectors wrap together)
        int c = lane < HV ? xw0 - L::A + lane * VEC : xw0 + wcols + (lane - HV) * VEC;
        c += c < 0 ? nx : 0;
        c -= c >= nx ? nx : 0;
        halo_col = c;
        halo_slot = lane < HV ? lane * VEC : L::A + wcols + (lane - HV) * VEC;
    }
    __device__ __forceinline__ void load(const T *__restrict__ row, V &raw, V &hal) const {
        raw = active ? vec_load<T, VEC>(row + x) : vec_zero<T, VEC>();
        if (has_halo) hal = vec_load<T, VEC>(row + halo_col);
    }
    // x-blur of the row whose vectors are (raw, hal); buf = the warp's shared-memory row of this parity
    __device__ __forceinline__ V blur(const TapsR<T, RX> &tx, const V &raw, const V &hal, T *buf) const {
        if (RX == 0) return raw;
        if (active) vec_store<T, VEC>(buf + L::A + lane * VEC, raw);      // (the slots past a partial strip hold the right halo)
        if (has_halo) vec_store<T, VEC>(buf + halo_slot, hal);
        __syncwarp();
        V out = vec_zero<T, VEC>();
        if (active) {
            T val[VEC + 2 * L::A];
#pragma unroll
            for (int j = 0; j < (VEC + 2 * L::A) / VEC; ++j) {
                const V w = vec_load<T, VEC>(buf + lane * VEC + j * VEC);
#pragma unroll
                for (int v = 0; v < VEC; ++v) val[j * VEC + v] = w.v[v];
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                T acc = T(0);
#pragma unroll
                for (int k = 0; k <= 2 * RX; ++k) acc += tx.t[k] * val[L::A + v + RX - k];
                out.v[v] = acc;
            }
        }
        return out;
    }
};

struct Fused2dGeom {
    int nx, nz, zc;
    long long n;
};

// cp.async (LDGSTS) staging of the operand rows of the NEXT output row: bytes in flight without registers
__device__ __forceinline__ void fused2d_cp16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void fused2d_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fused2d_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ int fused2d_wrap(int z, int nz) {
    z += z < 0 ? nz : 0;
    z -= z >= nz ? nz : 0;
    return z;
}

template <typename T, int RX, int RZ, int VEC>
__global__ void __launch_bounds__(32 * FUSED2D_WARPS, FUSED2D_MINB) fused2d_fwd_kernel(Fused2dGeom g, T wx, T wz, const LsmrScalars *__restrict__ S,
                                                                         TapsR<T, RX> tx, TapsR<T, RZ> tz, const T *__restrict__ vhat,
                                                                         T *__restrict__ u, double *__restrict__ part) {
    using V = Vec<T, VEC>;
    using L = Fused2dRow<T, VEC, RX>;
    __shared__ __align__(16) T srow[FUSED2D_WARPS][2][L::LEN];
    __shared__ __align__(16) T sop[FUSED2D_WARPS][2][5][L::W];            // operand rows of the next / current output row
    if (S->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xw0 = ((int)blockIdx.x * FUSED2D_WARPS + warp) * L::W;      // first column of the warp's strip
    const bool warp_in = xw0 < g.nx;
    const int z0 = (int)blockIdx.y * g.zc;
    const int z1 = min(g.nz, z0 + g.zc);
    const T inv_alpha = (T)S->inv_alpha, inv_beta = (T)S->inv_beta, malpha = (T)(-S->alpha), sa = (T)S->sqrt_alpha;
    T *u0 = u, *u1 = u + g.n, *u2 = u + 2 * g.n;
    double acc = 0.0;
    if (warp_in) {
        Fused2dLane<T, VEC, RX> ln;
        ln.init(xw0, g.nx, lane);
        const int x = ln.x;
        const bool active = ln.active;
        V ring[2 * RZ + 1];
#pragma unroll
        for (int i = 0; i <= 2 * RZ; ++i) ring[i] = vec_zero<T, VEC>();
        const int steps = (z1 - z0) + 2 * RZ;
        V raw_n = vec_zero<T, VEC>(), hal_n = vec_zero<T, VEC>();
        ln.load(vhat + (long long)fused2d_wrap(z0 - RZ, g.nz) * g.nx, raw_n, hal_n);
        T right_n = T(0);
        for (int j = 0; j < steps; ++j) {
            const V raw = raw_n, hal = hal_n;
            T right = right_n;
            if (j + 1 < steps) ln.load(vhat + (long long)fused2d_wrap(z0 - RZ + j + 1, g.nz) * g.nx, raw_n, hal_n);   // next input row
            // operand rows of the NEXT output row (z + 1) into the other stage: in flight during this step's work
            const int z = z0 + j - 2 * RZ;               // output row of this step (if j >= 2 RZ)
            if (j + 1 >= 2 * RZ && j + 1 < steps && active) {
                const long long in = (long long)(z + 1) * g.nx + x;
                T *st = &sop[warp][(j + 1) & 1][0][lane * VEC];
                fused2d_cp16(st, vhat + in);
                if (z + 2 < g.nz) fused2d_cp16(st + L::W, vhat + in + g.nx);
                fused2d_cp16(st + 2 * L::W, u0 + in);
                fused2d_cp16(st + 3 * L::W, u1 + in);
                fused2d_cp16(st + 4 * L::W, u2 + in);
                right_n = (lane == 31 && x + VEC < g.nx) ? vhat[in + VEC] : T(0);
            }
            fused2d_commit();
            const bool outp = j >= 2 * RZ;
            const long long i0 = (long long)z * g.nx + x;
            const V hx = ln.blur(tx, raw, hal, srow[warp][j & 1]);
#pragma unroll
            for (int i = 0; i < 2 * RZ; ++i) ring[i] = ring[i + 1];
            ring[2 * RZ] = hx;                           // ring[i] = x-blurred row z - RZ + i
            if (!outp) continue;
            fused2d_wait_prev();                         // the group committed one step ago (this row's operands) has landed
            V vr = vec_zero<T, VEC>(), vd = vec_zero<T, VEC>(), a0 = vr, a1 = vr, a2 = vr;
            if (active) {
                const T *st = &sop[warp][j & 1][0][lane * VEC];
                vr = vec_load<T, VEC>(st);
                if (z + 1 < g.nz) vd = vec_load<T, VEC>(st + L::W);
                a0 = vec_load<T, VEC>(st + 2 * L::W);
                a1 = vec_load<T, VEC>(st + 3 * L::W);
                a2 = vec_load<T, VEC>(st + 4 * L::W);
            }
            const T nb = __shfl_down_sync(0xffffffffu, vr.v[0], 1);      // first value of the lane to the right
            if (lane != 31) right = (x + VEC < g.nx) ? nb : T(0);
            if (active) {
                V vc;
#pragma unroll
                for (int v = 0; v < VEC; ++v) vc.v[v] = vr.v[v] * inv_alpha;
                const T rgt = right * inv_alpha;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    T av = T(0);
#pragma unroll
                    for (int k = 0; k <= 2 * RZ; ++k) av += tz.t[k] * ring[2 * RZ - k].v[v];      // row z - (k - RZ)
                    T un = (a0.v[v] * inv_beta) * malpha + av * inv_alpha;
                    a0.v[v] = un;
                    acc += (double)un * (double)un;
                    const T hi = (v + 1 < VEC) ? vc.v[(v + 1) % VEC] : rgt;
                    const T dx = wx * hi + (-wx) * vc.v[v];
                    un = (a1.v[v] * inv_beta) * malpha + sa * dx;
                    a1.v[v] = un;
                    acc += (double)un * (double)un;
                    const T hz = (z + 1 < g.nz) ? vd.v[v] * inv_alpha : T(0);
                    const T dz = wz * hz + (-wz) * vc.v[v];
                    un = (a2.v[v] * inv_beta) * malpha + sa * dz;
                    a2.v[v] = un;
                    acc += (double)un * (double)un;
                }
                vec_store<T, VEC>(u0 + i0, a0);
                vec_store<T, VEC>(u1 + i0, a1);
                vec_store<T, VEC>(u2 + i0, a2);
            }
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[(long long)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

template <typename T, int RX, int RZ, int VEC>
__global__ void __launch_bounds__(32 * FUSED2D_WARPS, FUSED2D_MINB) fused2d_adj_kernel(Fused2dGeom g, T wx, T wz, const LsmrScalars *__restrict__ S,
                                                                         TapsR<T, RX> tx, TapsR<T, RZ> tz, const T *__restrict__ u,
                                                                         T *__restrict__ vhat, double *__restrict__ part, int first) {
    using V = Vec<T, VEC>;
    using L = Fused2dRow<T, VEC, RX>;
    __shared__ __align__(16) T srow[FUSED2D_WARPS][2][L::LEN];
    __shared__ __align__(16) T sop[FUSED2D_WARPS][2][3][L::W];            // operand rows of the next / current output row
    if (S->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xw0 = ((int)blockIdx.x * FUSED2D_WARPS + warp) * L::W;
    const bool warp_in = xw0 < g.nx;
    const int z0 = (int)blockIdx.y * g.zc;
    const int z1 = min(g.nz, z0 + g.zc);
    const T inv_alpha = (T)S->inv_alpha, inv_beta = (T)S->inv_beta, mbeta = (T)(-S->beta), sa = (T)S->sqrt_alpha;
    const T *u0 = u, *u1 = u + g.n, *u2 = u + 2 * g.n;
    double acc = 0.0;
    if (warp_in) {
        Fused2dLane<T, VEC, RX> ln;
        ln.init(xw0, g.nx, lane);
        const int x = ln.x;
        const bool active = ln.active;
        V ring[2 * RZ + 1];
#pragma unroll
        for (int i = 0; i <= 2 * RZ; ++i) ring[i] = vec_zero<T, VEC>();
        V u2_prev = vec_zero<T, VEC>();                  // u2 row z - 1 (zero above the first row: Dz^T boundary)
        if (active && z0 > 0) u2_prev = vec_load<T, VEC>(u2 + (long long)(z0 - 1) * g.nx + x);
        const int steps = (z1 - z0) + 2 * RZ;
        V raw_n = vec_zero<T, VEC>(), hal_n = vec_zero<T, VEC>();
        ln.load(u0 + (long long)fused2d_wrap(z0 - RZ, g.nz) * g.nx, raw_n, hal_n);
        T left_n = T(0);
        for (int j = 0; j < steps; ++j) {
            const V raw = raw_n, hal = hal_n;
            T left = left_n;
            if (j + 1 < steps) ln.load(u0 + (long long)fused2d_wrap(z0 - RZ + j + 1, g.nz) * g.nx, raw_n, hal_n);
            const int z = z0 + j - 2 * RZ;
            if (j + 1 >= 2 * RZ && j + 1 < steps && active) {        // operand rows of the next output row
                const long long in = (long long)(z + 1) * g.nx + x;
                T *st = &sop[warp][(j + 1) & 1][0][lane * VEC];
                fused2d_cp16(st, u1 + in);
                fused2d_cp16(st + L::W, u2 + in);
                if (!first) fused2d_cp16(st + 2 * L::W, vhat + in);
                left_n = (lane == 0 && x > 0) ? u1[in - 1] : T(0);
            }
            fused2d_commit();
            const bool outp = j >= 2 * RZ;
            const long long i0 = (long long)z * g.nx + x;
            const V hx = ln.blur(tx, raw, hal, srow[warp][j & 1]);
#pragma unroll
            for (int i = 0; i < 2 * RZ; ++i) ring[i] = ring[i + 1];
            ring[2 * RZ] = hx;
            if (!outp) continue;
            fused2d_wait_prev();
            V b1 = vec_zero<T, VEC>(), b2 = b1, vv = b1;
            if (active) {
                const T *st = &sop[warp][j & 1][0][lane * VEC];
                b1 = vec_load<T, VEC>(st);
                b2 = vec_load<T, VEC>(st + L::W);
                if (!first) vv = vec_load<T, VEC>(st + 2 * L::W);
            }
            const T nb = __shfl_up_sync(0xffffffffu, b1.v[VEC - 1], 1);   // last value of the lane to the left
            if (lane != 0) left = nb;
            if (active) {
                const T lft = left * inv_beta;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    T av = T(0);
#pragma unroll
                    for (int k = 0; k <= 2 * RZ; ++k) av += tz.t[k] * ring[2 * RZ - k].v[v];
                    T r = av * inv_beta;
                    const T lo = (v == 0) ? lft : b1.v[(v + VEC - 1) % VEC] * inv_beta;
                    T div = wx * lo + (-wx) * (b1.v[v] * inv_beta);
                    const T loz = (z > 0) ? u2_prev.v[v] * inv_beta : T(0);
                    div = div + (wz * loz + (-wz) * (b2.v[v] * inv_beta));
                    r = r + sa * div;
                    const T vn = first ? r : (vv.v[v] * inv_alpha) * mbeta + r;
                    vv.v[v] = vn;
                    acc += (double)vn * (double)vn;
                }
                vec_store<T, VEC>(vhat + i0, vv);
            }
            u2_prev = b2;
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[(long long)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// when the fused 2-D kernels apply: 2-D, A = separable blur, B = grad, no slab mode, enough rows to amortise the ring
// warm-up (or forced by the "lsmr_fuse2d" tuning knob: 1 = on whenever possible, 2 = off)
static bool fused2d_ok(const nsol_lsmr_plan *pl, int b_op) {
    const GridView &gv = pl->gv;
    if (gv.dim != 2 || pl->slab || pl->desc.a_op != NSOL_A_BLUR || b_op != NSOL_B_GRAD || !fastv_ok(pl)) return false;
    if (pl->ctx->lsmr_fuse2d == 2) return false;
    const int rz = pl->desc.radius[0], rx = pl->desc.radius[1];
    if (rz != rx || rz < 1) return false;                    // instantiated for isotropic masks only
    if (gv.nz < 2 * rz + 1) return false;
    if (pl->ctx->lsmr_fuse2d == 1 || pl->ctx->lsmr_fuse2d == 3) return true;      // 3: forced, first-generation kernels
    // from 16 MB per vector (1536^2 float64: 86.9 vs 107.7 us per iteration; 1024^2 float32: 42.5 vs 39.3 us): below,
    // the row-mapped kernels have more threads in flight
    return gv.n * (long long)pl->esz >= (1ll << 24);
}

static int fused2d_rows_per_chunk(const nsol_lsmr_plan *pl, int vec) {
    const GridView &gv = pl->gv;
    const int strips = (gv.nx + 32 * vec * FUSED2D_WARPS - 1) / (32 * vec * FUSED2D_WARPS);
    // aim at >= 2 blocks of 128 threads per SM and warp-strip; never fewer than 16 rows per chunk
    int zc = 64;
    while (zc > 16 && (long long)strips * ((gv.nz + zc - 1) / zc) < (long long)pl->ctx->sm_count * 8) zc /= 2;
    if ((pl->ctx->lsmr_fuse2d == 1 || pl->ctx->lsmr_fuse2d == 3) && gv.nz < 64) zc = gv.nz < 8 ? gv.nz : 8;     // tests: several chunks on small images
    return zc;
}

template <typename T>
static int fused2d_launch(nsol_lsmr_plan *pl, bool forward, int first, cudaStream_t s, int *nparts) {
    constexpr int VEC = FastvCfg<T>::VEC;
    const GridView &gv = pl->gv;
    Fused2dGeom g;
    g.nx = gv.nx;
    g.nz = gv.nz;
    g.n = gv.n;
    g.zc = fused2d_rows_per_chunk(pl, VEC);
    const dim3 grid((gv.nx + 32 * VEC * FUSED2D_WARPS - 1) / (32 * VEC * FUSED2D_WARPS), (gv.nz + g.zc - 1) / g.zc, 1);
    const T wx = (T)gv.w[0], wz = (T)gv.w[1];
    const int r = pl->desc.radius[0];
    if (forward) {
        FASTV_SWITCH_R(r, (fused2d_fwd_kernel<T, (R < 1 ? 1 : R), (R < 1 ? 1 : R), VEC><<<grid, 32 * FUSED2D_WARPS, 0, s>>>(
                              g, wx, wz, pl->S, lsq_taps_r<T, (R < 1 ? 1 : R)>(pl, 1), lsq_taps_r<T, (R < 1 ? 1 : R)>(pl, 0), (const T *)pl->v,
                              (T *)pl->u, pl->part)));
    } else {
        FASTV_SWITCH_R(r, (fused2d_adj_kernel<T, (R < 1 ? 1 : R), (R < 1 ? 1 : R), VEC><<<grid, 32 * FUSED2D_WARPS, 0, s>>>(
                              g, wx, wz, pl->S, lsq_taps_r<T, (R < 1 ? 1 : R)>(pl, 1), lsq_taps_r<T, (R < 1 ? 1 : R)>(pl, 0), (const T *)pl->u,
                              (T *)pl->v, pl->part, first)));
    }
    *nparts = (int)(grid.x * grid.y);
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

// =====================================================================================================
// Vector versions of the per-solve / per-outer-iteration kernels (right-hand side, start vectors, clip, ADMM shrink):
// once per 10 LSMR iterations, but 21 words per pixel together -- as scalar grid-stride kernels with a 64-bit
// division per element they cost about two LSMR iterations.
// =====================================================================================================
// u = [b; sqrt_alpha * b_reg], partial ||u||^2.  nvec = n / VEC vectors per block of u.
template <typename T, int VEC>
__global__ void __launch_bounds__(LSMR_THREADS) fastv_rhs_kernel(long long nvec, int rows_b, const T *__restrict__ b, const T *__restrict__ breg,
                                                                 T sqrt_alpha, T *__restrict__ u, double *__restrict__ part,
                                                                 const double *__restrict__ sa_dev = nullptr) {
    using V = Vec<T, VEC>;
    if (sa_dev) sqrt_alpha = (T)*sa_dev;      // weight computed on the device (graph-replayed primal-dual deconvolution)
    const long long total = nvec * (1 + rows_b);
    double acc = 0.0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
        V v;
        if (j < nvec) v = vec_load<T, VEC>(b + j * VEC);
        else if (breg) {
            v = vec_load<T, VEC>(breg + (j - nvec) * VEC);
#pragma unroll
            for (int e = 0; e < VEC; ++e) v.v[e] = sqrt_alpha * v.v[e];
        } else v = vec_zero<T, VEC>();
        vec_store<T, VEC>(u + j * VEC, v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc += (double)v.v[e] * (double)v.v[e];
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// h = v, hbar = 0, x = 0   (lsmr.py:277-278, 253)
template <typename T, int VEC>
__global__ void __launch_bounds__(LSMR_THREADS) fastv_init_vectors_kernel(long long nvec, const LsmrScalars *__restrict__ S, const T *__restrict__ vhat,
                                                                          T *__restrict__ h, T *__restrict__ hbar, T *__restrict__ x) {
    using V = Vec<T, VEC>;
    const T inv_alpha = (T)S->inv_alpha;
    const V zero = vec_zero<T, VEC>();
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nvec; j += (long long)gridDim.x * blockDim.x) {
        V v = vec_load<T, VEC>(vhat + j * VEC);
#pragma unroll
        for (int e = 0; e < VEC; ++e) v.v[e] = v.v[e] * inv_alpha;
        vec_store<T, VEC>(h + j * VEC, v);
        vec_store<T, VEC>(hbar + j * VEC, zero);
        vec_store<T, VEC>(x + j * VEC, zero);
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(LSMR_THREADS) fastv_clip_kernel(long long nvec, const T *__restrict__ in, T *__restrict__ out, double lo, double hi) {
    using V = Vec<T, VEC>;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nvec; j += (long long)gridDim.x * blockDim.x) {
        V v = vec_load<T, VEC>(in + j * VEC);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            double d = (double)v.v[e];
            d = d < lo ? lo : (d > hi ? hi : d);     // np.clip
            v.v[e] = (T)d;
        }
        vec_store<T, VEC>(out + j * VEC, v);
    }
}

// ADMM: t = B x + w_in ; v = shrink_iso(t, ell) ; w = t - v ; b_reg = v - w      (admm_linear_solver.py:208-216, 222, 239-253)
// plain: v = B x, w = 0, b_reg = v (start of the run, :171-172).  Row-mapped like the forward kernel.
template <typename T, int VEC>
__global__ void __launch_bounds__(FAST_TH) fastv_shrink_kernel(FastvGeom<T> g, const T *__restrict__ x, const T *w_in, T ell, T *v_out, T *w_out,
                                                               T *breg_out, const T *__restrict__ x_hi, int plain, const T *__restrict__ c) {
    using V = Vec<T, VEC>;
    const int x0 = (int)(blockIdx.x * FAST_TH + threadIdx.x) * VEC;
    if (x0 >= g.nx) return;
    const int y = (int)blockIdx.y, z = (int)blockIdx.z;
    const long long plane = (long long)g.nx * g.ny;
    const long long i = (long long)z * plane + (long long)y * g.nx + x0;
    const V xc = vec_load<T, VEC>(x + i);
    V t[3], ss;
    // component 0: along x
    {
        const T right = (x0 + VEC < g.nx) ? x[i + VEC] : T(0);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const T hi = (e + 1 < VEC) ? xc.v[(e + 1) % VEC] : right;
            t[0].v[e] = g.wx * hi + (-g.wx) * xc.v[e];
        }
    }
    int nc = 1;
    if (g.dim == 3) {   // component 1: along y
        V hv = vec_zero<T, VEC>();
        if (y + 1 < g.ny) hv = vec_load<T, VEC>(x + i + g.nx);
#pragma unroll
        for (int e = 0; e < VEC; ++e) t[1].v[e] = g.wy * hv.v[e] + (-g.wy) * xc.v[e];
        nc = 2;
    }
    if (g.dim >= 2) {   // last component: along z
        V hv = vec_zero<T, VEC>();
        if (z + 1 < g.nz) hv = vec_load<T, VEC>(x + i + plane);
        else if (g.slab && g.grad_hi) hv = vec_load<T, VEC>(x_hi + (long long)y * g.nx + x0);
        V tz;
#pragma unroll
        for (int e = 0; e < VEC; ++e) tz.v[e] = g.wz * hv.v[e] + (-g.wz) * xc.v[e];
        if (nc == 2) t[2] = tz;     // static register slots (no dynamically indexed local array)
        else t[1] = tz;
        nc += 1;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < nc) {
            if (w_in) {
                const V wv = vec_load<T, VEC>(w_in + (long long)k * g.n + i);
#pragma unroll
                for (int e = 0; e < VEC; ++e) t[k].v[e] = t[k].v[e] + wv.v[e];
            }
            if (c) {       // the solver's own b_reg: t = B x + w - c   (admm_linear_solver.py:208)
                const V cv = vec_load<T, VEC>(c + (long long)k * g.n + i);
#pragma unroll
                for (int e = 0; e < VEC; ++e) t[k].v[e] = t[k].v[e] - cv.v[e];
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) ss.v[e] = (k == 0) ? t[k].v[e] * t[k].v[e] : ss.v[e] + t[k].v[e] * t[k].v[e];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < nc) {
            V vk, wk, bk;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (plain) {
                    vk.v[e] = t[k].v[e];
                    wk.v[e] = T(0);
                    bk.v[e] = t[k].v[e];
                } else {
                    const T nrm = sqrt_t(ss.v[e]);
                    const bool on = nrm > ell;
                    const T soft = max_t(nrm - ell, T(0));     // |n| = n >= 0, sign(n) = 1 where n > ell
                    vk.v[e] = on ? soft * t[k].v[e] / nrm : T(0);
                    wk.v[e] = t[k].v[e] - vk.v[e];
                    bk.v[e] = vk.v[e] - wk.v[e];
                }
            }
            if (c && breg_out) {       // b_reg of the next solve = v - w + c   (:222)
                const V cv = vec_load<T, VEC>(c + (long long)k * g.n + i);
#pragma unroll
                for (int e = 0; e < VEC; ++e) bk.v[e] = bk.v[e] + cv.v[e];
            }
            vec_store<T, VEC>(v_out + (long long)k * g.n + i, vk);
            vec_store<T, VEC>(w_out + (long long)k * g.n + i, wk);
            if (breg_out) vec_store<T, VEC>(breg_out + (long long)k * g.n + i, bk);
        }
    }
}

// =====================================================================================================
// Vector versions of the primal-dual deconvolution kernels (pdd_dual / pdd_arg / pdd_relax in lsmr_kernels.cu):
// same arithmetic per element (the divisions of the reference are kept), row-mapped, 16 bytes of x per thread.
// =====================================================================================================
// p <- prox_g*(p + sigma grad(xbar))   (primal_dual_solver.py:242-243; proximal_operators.py:139-140, 157-159)
template <typename T, int VEC>
__global__ void __launch_bounds__(FAST_TH) fastv_pdd_dual_kernel(FastvGeom<T> g, const T *__restrict__ xbar, T *__restrict__ p,
                                                                 const double *__restrict__ sched, const int *__restrict__ it, int reg) {
    using V = Vec<T, VEC>;
    const double *srow = sched + (long long)(*it) * 8;      // step sizes of the current iteration (device table)
    const T sigma = (T)srow[0], den = (T)srow[4];
    const int x0 = (int)(blockIdx.x * FAST_TH + threadIdx.x) * VEC;
    if (x0 >= g.nx) return;
    const int y = (int)blockIdx.y, z = (int)blockIdx.z;
    const long long plane = (long long)g.nx * g.ny;
    const long long i = (long long)z * plane + (long long)y * g.nx + x0;
    const V xc = vec_load<T, VEC>(xbar + i);
    auto finish = [&](int k, const V &gk) {
        V q = vec_load<T, VEC>(p + (long long)k * g.n + i);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            T v = q.v[e] + sigma * gk.v[e];
            if (reg != NSOL_REG_TV) v = v / den;
            if (reg != NSOL_REG_TK1) v = v / max_t(T(1), abs_t(v));
            q.v[e] = v;
        }
        vec_store<T, VEC>(p + (long long)k * g.n + i, q);
    };
    {
        const T right = (x0 + VEC < g.nx) ? xbar[i + VEC] : T(0);
        V gk;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const T hi = (e + 1 < VEC) ? xc.v[(e + 1) % VEC] : right;
            gk.v[e] = g.wx * hi + (-g.wx) * xc.v[e];
        }
        finish(0, gk);
    }
    if (g.dim == 3) {
        V hv = vec_zero<T, VEC>(), gk;
        if (y + 1 < g.ny) hv = vec_load<T, VEC>(xbar + i + g.nx);
#pragma unroll
        for (int e = 0; e < VEC; ++e) gk.v[e] = g.wy * hv.v[e] + (-g.wy) * xc.v[e];
        finish(1, gk);
    }
    if (g.dim >= 2) {
        V hv = vec_zero<T, VEC>(), gk;
        if (z + 1 < g.nz) hv = vec_load<T, VEC>(xbar + i + plane);
#pragma unroll
        for (int e = 0; e < VEC; ++e) gk.v[e] = g.wz * hv.v[e] + (-g.wz) * xc.v[e];
        finish(g.dim - 1, gk);
    }
}

// b_reg <- (x - tau grad_adj(p)) / prox_scale   (primal_dual_solver.py:246; tikhonov b_reg / x_scale)
template <typename T, int VEC>
__global__ void __launch_bounds__(FAST_TH) fastv_pdd_arg_kernel(FastvGeom<T> g, const T *__restrict__ x, const T *__restrict__ p,
                                                                const double *__restrict__ sched, const int *__restrict__ it, T prox_scale,
                                                                T *__restrict__ breg) {
    using V = Vec<T, VEC>;
    const T tau = (T)sched[(long long)(*it) * 8 + 1];
    const int x0 = (int)(blockIdx.x * FAST_TH + threadIdx.x) * VEC;
    if (x0 >= g.nx) return;
    const int y = (int)blockIdx.y, z = (int)blockIdx.z;
    const long long plane = (long long)g.nx * g.ny;
    const long long i = (long long)z * plane + (long long)y * g.nx + x0;
    V div;
    {
        const T *pk = p;
        const V pv = vec_load<T, VEC>(pk + i);
        const T left = (x0 > 0) ? pk[i - 1] : T(0);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const T lo = (e == 0) ? left : pv.v[(e + VEC - 1) % VEC];
            div.v[e] = g.wx * lo + (-g.wx) * pv.v[e];
        }
    }
    if (g.dim == 3) {
        const T *pk = p + g.n;
        const V pv = vec_load<T, VEC>(pk + i);
        V lv = vec_zero<T, VEC>();
        if (y > 0) lv = vec_load<T, VEC>(pk + i - g.nx);
#pragma unroll
        for (int e = 0; e < VEC; ++e) div.v[e] = div.v[e] + (g.wy * lv.v[e] + (-g.wy) * pv.v[e]);
    }
    if (g.dim >= 2) {
        const T *pk = p + (long long)(g.dim - 1) * g.n;
        const V pv = vec_load<T, VEC>(pk + i);
        V lv = vec_zero<T, VEC>();
        if (z > 0) lv = vec_load<T, VEC>(pk + i - plane);
#pragma unroll
        for (int e = 0; e < VEC; ++e) div.v[e] = div.v[e] + (g.wz * lv.v[e] + (-g.wz) * pv.v[e]);
    }
    const V xv = vec_load<T, VEC>(x + i);
    V out;
#pragma unroll
    for (int e = 0; e < VEC; ++e) out.v[e] = (xv.v[e] - tau * div.v[e]) / prox_scale;
    vec_store<T, VEC>(breg + i, out);
}

// x+ = y * prox_scale ; xbar = x+ + theta (x+ - x) ; x = x+   (solver.py:117-118; primal_dual_solver.py:253)
template <typename T, int VEC>
__global__ void __launch_bounds__(LSMR_THREADS) fastv_pdd_relax_kernel(long long nvec, const T *__restrict__ yv, T prox_scale,
                                                                       const double *__restrict__ sched, const int *__restrict__ it, T *__restrict__ x,
                                                                       T *__restrict__ xbar) {
    using V = Vec<T, VEC>;
    const T theta = (T)sched[(long long)(*it) * 8 + 3];
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nvec; j += (long long)gridDim.x * blockDim.x) {
        const V a = vec_load<T, VEC>(yv + j * VEC), xo = vec_load<T, VEC>(x + j * VEC);
        V xn, xb;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            xn.v[e] = a.v[e] * prox_scale;
            xb.v[e] = xn.v[e] + theta * (xn.v[e] - xo.v[e]);
        }
        vec_store<T, VEC>(xbar + j * VEC, xb);
        vec_store<T, VEC>(x + j * VEC, xn);
    }
}
