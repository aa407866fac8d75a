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
template <typename T, int R, int VEC>
__global__ void __launch_bounds__(FAST_TH) fastv_blur_pass_kernel(FastvGeom<T> g, TapsR<T, R> tp, int kaxis, const T *__restrict__ in,
                                                                  T *__restrict__ out, const T *__restrict__ halo_lo,
                                                                  const T *__restrict__ halo_hi) {
    using V = Vec<T, VEC>;
    const int x = (int)(blockIdx.x * FAST_TH + threadIdx.x) * VEC;
    if (x >= g.nx) return;
    // the grid axis of the blurred direction counts groups of FASTV_ROWS positions
    const int y = kaxis == 1 ? 0 : (int)blockIdx.y, z = kaxis == 2 ? 0 : (int)blockIdx.z;
    const int pos0 = (kaxis == 1 ? (int)blockIdx.y : (int)blockIdx.z) * FASTV_ROWS;
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
__global__ void __launch_bounds__(FAST_TH) fastv_fwd_kernel(FastvGeom<T> g, const LsmrScalars *__restrict__ S, TapsR<T, R> tx,
                                                            const T *__restrict__ src, const T *__restrict__ vhat, T *__restrict__ u,
                                                            double *__restrict__ part, const T *__restrict__ v_hi) {
    using V = Vec<T, VEC>;
    if (S->done) return;
    const int x = (int)(blockIdx.x * FAST_TH + threadIdx.x) * VEC;
    const int y = (int)blockIdx.y, z = (int)blockIdx.z;
    double acc = 0.0;
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
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
}

// vhat <- (vhat * inv_alpha) * (-beta) + (A^T u0 + sqrt_alpha B^T u1..), u = uhat * inv_beta, partial ||v||^2 (lsmr.py:342-344)
template <typename T, int R, int VEC>
__global__ void __launch_bounds__(FAST_TH) fastv_adj_kernel(FastvGeom<T> g, const LsmrScalars *__restrict__ S, TapsR<T, R> tx,
                                                            const T *__restrict__ src, const T *__restrict__ u, T *__restrict__ vhat,
                                                            double *__restrict__ part, int first, const T *__restrict__ uz_lo) {
    using V = Vec<T, VEC>;
    if (S->done) return;
    const int x = (int)(blockIdx.x * FAST_TH + threadIdx.x) * VEC;
    const int y = (int)blockIdx.y, z = (int)blockIdx.z;
    double acc = 0.0;
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
