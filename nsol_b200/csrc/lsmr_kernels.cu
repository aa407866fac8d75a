// lsmr_kernels.cu -- LSMR on the stacked system [A; sqrt(alpha) B] x = [b; sqrt(alpha) b_reg]
// and the ADMM TV-L2 loop on top of it, entirely device-resident.
//
// Replaces (reference file:line)
//   TikhonovLinearSolver._run, lsmr/linear branch   nsol/tikhonov_linear_solver.py:120-158
//   _get_augmented_linear_system / _A_augmented(_adj) nsol/tikhonov_linear_solver.py:226-274
//   scipy.sparse.linalg.lsmr (scipy 1.18.1, Fong & Saunders 2011), cold start, atol=btol=0,
//     conlim=1e8                                      scipy/sparse/linalg/_isolve/lsmr.py:239-479
//   ADMMLinearSolver._run / _perform_ADMM_iteration / _prox_g  nsol/admm_linear_solver.py:165-253
//
// Layout: u has (1 + rows_B) SoA blocks of N (block 0 = data rows, blocks 1.. = rows of B);
// v, h, hbar, x have N elements.  u and v are stored UN-normalised together with 1/beta and
// 1/alpha; the scaling is applied by the consumers (no separate scaling pass).  All scalar
// recurrences (Givens rotations, norm estimates, stopping tests) run in float64 in a
// one-block kernel between the vector kernels, so an LSMR solve -- and a whole ADMM run --
// needs no host synchronisation.  Norms: per-block partial sums reduced in a fixed order
// (deterministic, float64 accumulation).
#include <vector>

#include "common.cuh"

#define LSMR_THREADS 256

struct LsmrScalars {
    double alpha, beta, inv_alpha, inv_beta;
    double zetabar, alphabar, rho, rhobar, cbar, sbar;
    double betadd, betad, rhodold, tautildeold, thetatilde, zeta, d;
    double normA2, maxrbar, minrbar, normb, normr, normar, normA, condA, normx;
    double c_hbar, c_x, c_h;   // coefficients of the vector update
    double sqrt_alpha;         // weight of the B rows
    int itn, istop, done, maxiter;
};

// ---------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    __syncthreads();   // protect warp_part against a previous call
    if (lane == 0) warp_part[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < nwarps ? warp_part[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    return r;   // valid in thread 0
}

// sum of `count` partials in a fixed order (deterministic); result valid in thread 0 (one block).
// Four independent accumulators per thread keep several loads in flight.
__device__ __forceinline__ double reduce_partials(const double *part, int count) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int step = blockDim.x;
    int i = threadIdx.x;
    for (; i + 3 * step < count; i += 4 * step) {
        a0 += part[i];
        a1 += part[i + step];
        a2 += part[i + 2 * step];
        a3 += part[i + 3 * step];
    }
    for (; i < count; i += step) a0 += part[i];
    return block_sum((a0 + a1) + (a2 + a3));
}

// ---------------------------------------------------------------------------
// stable Givens rotation (S.-C. Choi's SymOrtho; scipy lsqr.py:62-94)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double sign_d(double a) { return a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : 0.0); }

__device__ __forceinline__ void sym_ortho(double a, double b, double &c, double &s, double &r) {
    if (b == 0.0) {
        c = sign_d(a); s = 0.0; r = fabs(a);
    } else if (a == 0.0) {
        c = 0.0; s = sign_d(b); r = fabs(b);
    } else if (fabs(b) > fabs(a)) {
        const double tau = a / b;
        s = sign_d(b) / sqrt(1.0 + tau * tau);
        c = s * tau;
        r = b / s;
    } else {
        const double tau = b / a;
        c = sign_d(a) / sqrt(1.0 + tau * tau);
        s = c * tau;
        r = a / c;
    }
}

// ---------------------------------------------------------------------------
// scalar kernels (one block)
// ---------------------------------------------------------------------------
// The scalar steps as device functions on a LsmrScalars block (global memory for the multi-kernel
// path, shared memory for the cooperative kernel).  `ss` is the freshly reduced sum of squares.

// after the right-hand side has been written into u: beta = ||b||
__device__ __forceinline__ void scal_init_beta(LsmrScalars *S, double ss, double sqrt_alpha, int maxiter) {
    const double beta = sqrt(ss);
    S->normb = beta;
    S->beta = beta;
    S->inv_beta = beta > 0.0 ? 1.0 / beta : 0.0;
    S->sqrt_alpha = sqrt_alpha;
    S->maxiter = maxiter;
    S->itn = 0;
    S->istop = 0;
    S->done = 0;
    S->alpha = 0.0;
    S->inv_alpha = 0.0;
}

// after v = A^T u: alpha = ||v|| and the initial values of lsmr.py:262-313
__device__ __forceinline__ void scal_init_alpha(LsmrScalars *S, double ss) {
    const double alpha = S->beta > 0.0 ? sqrt(ss) : 0.0;
    S->alpha = alpha;
    S->inv_alpha = alpha > 0.0 ? 1.0 / alpha : 0.0;
    S->zetabar = alpha * S->beta;
    S->alphabar = alpha;
    S->rho = 1.0; S->rhobar = 1.0; S->cbar = 1.0; S->sbar = 0.0;
    S->betadd = S->beta; S->betad = 0.0; S->rhodold = 1.0; S->tautildeold = 0.0;
    S->thetatilde = 0.0; S->zeta = 0.0; S->d = 0.0;
    S->normA2 = alpha * alpha; S->maxrbar = 0.0; S->minrbar = 1e100;
    S->normA = sqrt(S->normA2); S->condA = 1.0; S->normx = 0.0;
    S->normr = S->beta;
    S->normar = alpha * S->beta;
    // lsmr.py:307-315: nothing to do if A^T b = 0 or b = 0 (x stays 0)
    if (S->normar == 0.0 || S->normb == 0.0 || S->maxiter <= 0) S->done = 1;
}

// beta_{k+1} = ||u||  (lsmr.py:338-340)
__device__ __forceinline__ void scal_beta(LsmrScalars *S, double ss) {
    const double beta = sqrt(ss);
    S->beta = beta;
    S->inv_beta = beta > 0.0 ? 1.0 / beta : 0.0;
}

// alpha_{k+1} = ||v|| and the two plane rotations (lsmr.py:344-377, 379-412)
__device__ __forceinline__ void scal_alpha(LsmrScalars *S, double ss) {
    double alpha = S->alpha;
    const double beta = S->beta;
    if (beta > 0.0) {
        alpha = sqrt(ss);
        S->alpha = alpha;
        S->inv_alpha = alpha > 0.0 ? 1.0 / alpha : S->inv_alpha;
    }
    S->itn += 1;
    const double damp = 0.0;
    double chat, shat, alphahat;
    sym_ortho(S->alphabar, damp, chat, shat, alphahat);
    const double rhoold = S->rho;
    double c, s, rho;
    sym_ortho(alphahat, beta, c, s, rho);
    const double thetanew = s * alpha;
    S->alphabar = c * alpha;
    const double rhobarold = S->rhobar;
    const double zetaold = S->zeta;
    const double thetabar = S->sbar * rho;
    const double rhotemp = S->cbar * rho;
    double cbar, sbar, rhobar;
    sym_ortho(S->cbar * rho, thetanew, cbar, sbar, rhobar);
    const double zeta = cbar * S->zetabar;
    S->zetabar = -sbar * S->zetabar;
    S->cbar = cbar; S->sbar = sbar; S->rhobar = rhobar; S->rho = rho; S->zeta = zeta;
    // coefficients of lsmr.py:373-377
    S->c_hbar = -(thetabar * rho / (rhoold * rhobarold));
    S->c_x = zeta / (rho * rhobar);
    S->c_h = -(thetanew / rho);
    // estimate of ||r|| (lsmr.py:381-404)
    const double betaacute = chat * S->betadd;
    const double betacheck = -shat * S->betadd;
    const double betahat = c * betaacute;
    S->betadd = -s * betaacute;
    const double thetatildeold = S->thetatilde;
    double ctildeold, stildeold, rhotildeold;
    sym_ortho(S->rhodold, thetabar, ctildeold, stildeold, rhotildeold);
    S->thetatilde = stildeold * rhobar;
    S->rhodold = ctildeold * rhobar;
    S->betad = -stildeold * S->betad + ctildeold * betahat;
    S->tautildeold = (zetaold - thetatildeold * S->tautildeold) / rhotildeold;
    const double taud = (zeta - S->thetatilde * S->tautildeold) / S->rhodold;
    S->d = S->d + betacheck * betacheck;
    S->normr = sqrt(S->d + (S->betad - taud) * (S->betad - taud) + S->betadd * S->betadd);
    // estimate of ||A|| and cond(A) (lsmr.py:406-416)
    S->normA2 = S->normA2 + beta * beta;
    S->normA = sqrt(S->normA2);
    S->normA2 = S->normA2 + alpha * alpha;
    S->maxrbar = fmax(S->maxrbar, rhobarold);
    if (S->itn > 1) S->minrbar = fmin(S->minrbar, rhobarold);
    S->condA = fmax(S->maxrbar, rhotemp) / fmin(S->minrbar, rhotemp);
    S->normar = fabs(S->zetabar);
}

// ||x|| and the stopping rules with atol = btol = 0, conlim = 1e8 (lsmr.py:418-459)
__device__ __forceinline__ void scal_tests(LsmrScalars *S, double ss) {
    S->normx = sqrt(ss);
    const double normb = S->normb, normA = S->normA, normr = S->normr;
    const double test1 = normr / normb;
    const double test2 = (normA * normr) != 0.0 ? S->normar / (normA * normr) : INFINITY;
    const double test3 = 1.0 / S->condA;
    const double t1 = test1 / (1.0 + normA * S->normx / normb);
    const double rtol = 0.0;   // btol + atol * normA * normx / normb
    const double ctol = 1.0 / 1e8;
    int istop = 0;
    if (S->itn >= S->maxiter) istop = 7;
    if (1.0 + test3 <= 1.0) istop = 6;
    if (1.0 + test2 <= 1.0) istop = 5;
    if (1.0 + t1 <= 1.0) istop = 4;
    if (test3 <= ctol) istop = 3;
    if (test2 <= 0.0) istop = 2;
    if (test1 <= rtol) istop = 1;
    S->istop = istop;
    if (istop > 0) S->done = 1;
}

__global__ void lsmr_scalar_init_beta(LsmrScalars *S, const double *part, int count, double sqrt_alpha, int maxiter,
                                      const double *sa_dev = nullptr) {
    if (sa_dev) sqrt_alpha = *sa_dev;
    const double ss = reduce_partials(part, count);
    if (threadIdx.x == 0) scal_init_beta(S, ss, sqrt_alpha, maxiter);
}
__global__ void lsmr_scalar_init_alpha(LsmrScalars *S, const double *part, int count) {
    const double ss = reduce_partials(part, count);
    if (threadIdx.x == 0) scal_init_alpha(S, ss);
}
__global__ void lsmr_scalar_beta(LsmrScalars *S, const double *part, int count) {
    if (S->done) return;
    const double ss = reduce_partials(part, count);
    if (threadIdx.x == 0) scal_beta(S, ss);
}
__global__ void lsmr_scalar_alpha(LsmrScalars *S, const double *part, int count) {
    if (S->done) return;
    const double ss = reduce_partials(part, count);
    if (threadIdx.x == 0) scal_alpha(S, ss);
}
__global__ void lsmr_scalar_tests(LsmrScalars *S, const double *part, int count) {
    if (S->done) return;
    const double ss = reduce_partials(part, count);
    if (threadIdx.x == 0) scal_tests(S, ss);
}

// ---------------------------------------------------------------------------
// vector kernels
// ---------------------------------------------------------------------------
template <typename T>
struct LsqGeom {
    long long n;
    int nx, ny, nz, dim;
    long long stride[3];   // element stride of reference component k
    int extent[3];
    int axis[3];           // 0 = x, 1 = y, 2 = z (kernel axis of component k)
    T w[3];
    int b_op;              // nsol_bop
    // z-slab decomposition (one rank per GPU, nsol_lsmr_plan_slab): the volume is the local slab; planes
    // below plane 0 / above plane nz-1 live in halo buffers filled by the caller's neighbour exchange
    int slab;              // 0: whole volume (periodic wrap of the blur / zero boundary of the gradient in-kernel)
    int grad_lo, grad_hi;  // a neighbour exists below / above for the (non-periodic) gradient stencils
    int ghost;             // planes per halo buffer (>= blur radius along z, >= 1)
};

// plane offset of this thread's voxel inside its z-plane
template <typename T>
__device__ __forceinline__ long long lsq_plane_off(const LsqGeom<T> &g, const int idx[3]) {
    return (long long)idx[1] * g.nx + idx[0];
}

template <typename T>
__device__ __forceinline__ void lsq_decode(const LsqGeom<T> &g, long long r, int idx[3]) {
    if (g.n <= 0x7fffffffLL) {      // 32-bit division is ~5x cheaper than the 64-bit sequence
        const unsigned r32 = (unsigned)r, nx = (unsigned)g.nx, ny = (unsigned)g.ny;
        const unsigned t = r32 / nx;
        idx[0] = (int)(r32 - t * nx);
        idx[2] = (int)(t / ny);
        idx[1] = (int)(t - (unsigned)idx[2] * ny);
        return;
    }
    idx[0] = (int)(r % g.nx);
    const long long t = r / g.nx;
    idx[1] = (int)(t % g.ny);
    idx[2] = (int)(t / g.ny);
}

// u = [b; sqrt_alpha * b_reg], partial ||u||^2
template <typename T>
__global__ void lsmr_rhs_kernel(LsqGeom<T> g, int rows_b, const T *__restrict__ b, const T *__restrict__ breg, double sqrt_alpha,
                                T *__restrict__ u, double *__restrict__ part, const double *__restrict__ sa_dev = nullptr) {
    if (sa_dev) sqrt_alpha = *sa_dev;       // weight computed on the device (graph-replayed primal-dual deconvolution)
    const long long total = g.n * (1 + rows_b);
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        T v;
        if (i < g.n) v = b[i];
        else v = breg ? (T)sqrt_alpha * breg[i - g.n] : T(0);
        u[i] = v;
        acc += (double)v * (double)v;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// ---------------------------------------------------------------------------
// row-mapped vector kernels of the multi-kernel path: grid = (ceil(nx / FAST_TH), ny, nz), so a
// thread knows its (x, y, z) without any integer division, the separable blur wraps with a
// conditional add/subtract, and the LAST blur pass (along x) is fused into the consumer.
// ---------------------------------------------------------------------------
#define FAST_TH 256
#define FAST_MAX_TAPS 129

template <typename T>
struct TapArgs {
    T t[FAST_MAX_TAPS];
    int r;   // radius; -1: no blur (identity)
};

template <typename T>
__device__ __forceinline__ T fast_blur_line(const TapArgs<T> &tp, const T *line, int pos, int ext, long long st) {
    T acc = T(0);
    const int r = tp.r;
    if (r < ext) {
        for (int k = 0; k <= 2 * r; ++k) {
            int q = pos - (k - r);
            q += q < 0 ? ext : 0;
            q -= q >= ext ? ext : 0;
            acc += tp.t[k] * line[(long long)q * st];
        }
    } else {
        for (int k = 0; k <= 2 * r; ++k) {
            int q = (pos - (k - r)) % ext;
            if (q < 0) q += ext;
            acc += tp.t[k] * line[(long long)q * st];
        }
    }
    return acc;
}

// one separable pass along kernel axis kaxis (0 = x, 1 = y, 2 = z)
template <typename T>
__global__ void __launch_bounds__(FAST_TH) fast_blur_pass_kernel(LsqGeom<T> g, const __grid_constant__ TapArgs<T> tp, int kaxis,
                                                                 const T *__restrict__ in, T *__restrict__ out,
                                                                 const T *__restrict__ halo_lo, const T *__restrict__ halo_hi) {
    const int idx[3] = {(int)(blockIdx.x * FAST_TH + threadIdx.x), (int)blockIdx.y, (int)blockIdx.z};
    if (idx[0] >= g.nx) return;
    const long long i = ((long long)idx[2] * g.ny + idx[1]) * g.nx + idx[0];
    const long long st = kaxis == 0 ? 1 : (kaxis == 1 ? g.nx : (long long)g.nx * g.ny);
    const int ext = kaxis == 0 ? g.nx : (kaxis == 1 ? g.ny : g.nz);
    if (kaxis == 2 && g.slab) {
        // z-slab: the periodic neighbours are the halo planes (same tap order as fast_blur_line)
        const long long po = lsq_plane_off(g, idx);
        T acc = T(0);
        for (int k = 0; k <= 2 * tp.r; ++k) {
            const int q = idx[2] - (k - tp.r);
            const T *src = q < 0 ? halo_lo + (long long)(q + g.ghost) * st : (q >= g.nz ? halo_hi + (long long)(q - g.nz) * st : in + (long long)q * st);
            acc += tp.t[k] * src[po];
        }
        out[i] = acc;
        return;
    }
    out[i] = fast_blur_line(tp, in + (i - (long long)idx[kaxis] * st), idx[kaxis], ext, st);
}

// u <- (u * inv_beta) * (-alpha) + [A v; sqrt_alpha B v], v = vhat * inv_alpha, partial ||u||^2   (lsmr.py:336-338)
// src: vhat after the blur passes along z / y; the pass along x is applied here.
template <typename T>
__global__ void __launch_bounds__(FAST_TH) fast_fwd_kernel(LsqGeom<T> g, const LsmrScalars *__restrict__ S, const __grid_constant__ TapArgs<T> tx,
                                                           const T *__restrict__ src, const T *__restrict__ vhat, T *__restrict__ u,
                                                           double *__restrict__ part, const T *__restrict__ v_hi) {
    if (S->done) return;
    const int idx[3] = {(int)(blockIdx.x * FAST_TH + threadIdx.x), (int)blockIdx.y, (int)blockIdx.z};
    double acc = 0.0;
    if (idx[0] < g.nx) {
        const T inv_alpha = (T)S->inv_alpha, inv_beta = (T)S->inv_beta, malpha = (T)(-S->alpha), sa = (T)S->sqrt_alpha;
        const long long i = ((long long)idx[2] * g.ny + idx[1]) * g.nx + idx[0];
        const T av = (tx.r >= 0 ? fast_blur_line(tx, src + (i - idx[0]), idx[0], g.nx, 1) : src[i]) * inv_alpha;
        T un = (u[i] * inv_beta) * malpha + av;
        u[i] = un;
        acc += (double)un * (double)un;
        if (g.b_op == NSOL_B_GRAD) {
            const T vc = vhat[i] * inv_alpha;
            for (int k = 0; k < g.dim; ++k) {
                T hi = (idx[g.axis[k]] + 1 < g.extent[k]) ? vhat[i + g.stride[k]] * inv_alpha : T(0);
                if (g.slab && g.grad_hi && g.axis[k] == 2 && idx[2] + 1 == g.nz) hi = v_hi[lsq_plane_off(g, idx)] * inv_alpha;
                const T dk = g.w[k] * hi + (-g.w[k]) * vc;
                T *uk = u + (long long)(1 + k) * g.n;
                un = (uk[i] * inv_beta) * malpha + sa * dk;
                uk[i] = un;
                acc += (double)un * (double)un;
            }
        } else if (g.b_op == NSOL_B_IDENTITY) {
            T *uk = u + g.n;
            un = (uk[i] * inv_beta) * malpha + sa * (vhat[i] * inv_alpha);
            uk[i] = un;
            acc += (double)un * (double)un;
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
}

// vhat <- (vhat * inv_alpha) * (-beta) + (A^T u0 + sqrt_alpha B^T u1..), u = uhat * inv_beta, partial ||v||^2 (lsmr.py:342-344)
template <typename T>
__global__ void __launch_bounds__(FAST_TH) fast_adj_kernel(LsqGeom<T> g, const LsmrScalars *__restrict__ S, const __grid_constant__ TapArgs<T> tx,
                                                           const T *__restrict__ src, const T *__restrict__ u, T *__restrict__ vhat,
                                                           double *__restrict__ part, int first, const T *__restrict__ uz_lo) {
    if (S->done) return;
    const int idx[3] = {(int)(blockIdx.x * FAST_TH + threadIdx.x), (int)blockIdx.y, (int)blockIdx.z};
    double acc = 0.0;
    if (idx[0] < g.nx) {
        const T inv_alpha = (T)S->inv_alpha, inv_beta = (T)S->inv_beta, mbeta = (T)(-S->beta), sa = (T)S->sqrt_alpha;
        const long long i = ((long long)idx[2] * g.ny + idx[1]) * g.nx + idx[0];
        T r = (tx.r >= 0 ? fast_blur_line(tx, src + (i - idx[0]), idx[0], g.nx, 1) : src[i]) * inv_beta;
        if (g.b_op == NSOL_B_GRAD) {
            T div = T(0);
            for (int k = 0; k < g.dim; ++k) {
                const T *uk = u + (long long)(1 + k) * g.n;
                T lo = (idx[g.axis[k]] > 0) ? uk[i - g.stride[k]] * inv_beta : T(0);
                if (g.slab && g.grad_lo && g.axis[k] == 2 && idx[2] == 0) lo = uz_lo[lsq_plane_off(g, idx)] * inv_beta;
                const T dk = g.w[k] * lo + (-g.w[k]) * (uk[i] * inv_beta);
                div = (k == 0) ? dk : div + dk;
            }
            r = r + sa * div;
        } else if (g.b_op == NSOL_B_IDENTITY) {
            r = r + sa * (u[g.n + i] * inv_beta);
        }
        const T vn = first ? r : (vhat[i] * inv_alpha) * mbeta + r;
        vhat[i] = vn;
        acc += (double)vn * (double)vn;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
}

// h = v, hbar = 0, x = 0   (lsmr.py:277-278, 253)
template <typename T>
__global__ void lsmr_init_vectors_kernel(long long n, const LsmrScalars *__restrict__ S, const T *__restrict__ vhat, T *__restrict__ h,
                                         T *__restrict__ hbar, T *__restrict__ x) {
    const T inv_alpha = (T)S->inv_alpha;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        h[i] = vhat[i] * inv_alpha;
        hbar[i] = T(0);
        x[i] = T(0);
    }
}

// hbar = c_hbar*hbar + h;  x += c_x*hbar;  h = c_h*h + v;  partial ||x||^2   (lsmr.py:373-377, 421)
template <typename T>
__global__ void lsmr_update_kernel(long long n, const LsmrScalars *__restrict__ S, const T *__restrict__ vhat, T *__restrict__ h,
                                   T *__restrict__ hbar, T *__restrict__ x, double *__restrict__ part) {
    if (S->done) return;
    const T c_hbar = (T)S->c_hbar, c_x = (T)S->c_x, c_h = (T)S->c_h, inv_alpha = (T)S->inv_alpha;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T hv = h[i];
        const T hb = hbar[i] * c_hbar + hv;
        hbar[i] = hb;
        const T xv = x[i] + c_x * hb;
        x[i] = xv;
        h[i] = hv * c_h + vhat[i] * inv_alpha;
        acc += (double)xv * (double)xv;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

template <typename T>
__global__ void clip_kernel(long long n, const T *__restrict__ in, T *__restrict__ out, double lo, double hi) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = (double)in[i];
        v = v < lo ? lo : (v > hi ? hi : v);     // np.clip
        out[i] = (T)v;
    }
}

// ADMM: t = B x + w_in - c ; v = shrink_iso(t, ell) ; w = t - v      (admm_linear_solver.py:208-216, 239-253)
// optionally also breg_out = v - w + c  (the b_reg of the next Tikhonov solve, :222); c = the solver's own b_reg / x_scale
// (dim * N values, NULL = 0: admm_linear_solver.py:100)
template <typename T>
__global__ void admm_shrink_kernel(LsqGeom<T> g, const T *__restrict__ x, const T *__restrict__ w_in, T ell, T *__restrict__ v_out,
                                   T *__restrict__ w_out, T *__restrict__ breg_out, const T *__restrict__ x_hi = nullptr, int plain = 0,
                                   const T *__restrict__ c = nullptr) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += (long long)gridDim.x * blockDim.x) {
        int idx[3];
        lsq_decode(g, i, idx);
        const T xc = x[i];
        T t[3];
        T ss = T(0);
        for (int k = 0; k < g.dim; ++k) {
            T hi = (idx[g.axis[k]] + 1 < g.extent[k]) ? x[i + g.stride[k]] : T(0);
            if (g.slab && g.grad_hi && g.axis[k] == 2 && idx[2] + 1 == g.nz) hi = x_hi[lsq_plane_off(g, idx)];
            T tk = g.w[k] * hi + (-g.w[k]) * xc;
            if (w_in) tk = tk + w_in[(long long)k * g.n + i];
            if (c) tk = tk - c[(long long)k * g.n + i];
            t[k] = tk;
            ss = (k == 0) ? tk * tk : ss + tk * tk;
        }
        if (plain) {      // v = B x - c, w = 0, b_reg = v + c  (start of the ADMM run, admm_linear_solver.py:171-172, :222)
            for (int k = 0; k < g.dim; ++k) {
                v_out[(long long)k * g.n + i] = t[k];
                w_out[(long long)k * g.n + i] = T(0);
                breg_out[(long long)k * g.n + i] = c ? t[k] + c[(long long)k * g.n + i] : t[k];
            }
            continue;
        }
        const T nrm = sqrt_t(ss);
        const bool on = nrm > ell;
        const T soft = max_t(nrm - ell, T(0));     // |n| = n >= 0, sign(n) = 1 where n > ell
        for (int k = 0; k < g.dim; ++k) {
            const T vk = on ? soft * t[k] / nrm : T(0);
            const T wk = t[k] - vk;
            v_out[(long long)k * g.n + i] = vk;
            w_out[(long long)k * g.n + i] = wk;
            if (breg_out) breg_out[(long long)k * g.n + i] = c ? (vk - wk) + c[(long long)k * g.n + i] : vk - wk;
        }
    }
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
struct nsol_lsmr_plan {
    nsol_ctx *ctx = nullptr;
    cudaStream_t own_stream = nullptr;   // used when the caller passes the legacy default stream
    nsol_lsq_desc desc;
    nsol_grid grid;
    GridView gv;
    std::vector<double> taps[3];
    size_t esz = 8;
    int rows_b = 0;           // number of N-blocks of B rows
    int nblocks = 0;          // blocks of the flat vector kernels = number of their partials
    int row_blocks = 0;       // blocks (= partials) of the row-mapped kernels: ceil(nx/FAST_TH) * ny * nz
    void *u = nullptr, *v = nullptr, *h = nullptr, *hbar = nullptr, *x = nullptr;
    void *opbuf = nullptr, *optmp = nullptr;   // A v / A^T u and separable scratch
    void *breg = nullptr;                      // b_reg staging (rows_b * N)
    void *admm_v = nullptr, *admm_w = nullptr; // ADMM split variables (dim * N each)
    void *admm_c = nullptr;                    // the ADMM solver's own b_reg / x_scale (dim * N; NULL = 0), nsol_admm_set_b_reg_host
    void *bbuf = nullptr, *xbuf = nullptr;     // scaled observation / current solution (N)
    void *stage = nullptr;                     // float64 host-transfer staging
    size_t stage_bytes = 0;
    double *part = nullptr;
    LsmrScalars *S = nullptr;
    size_t bytes = 0;
    // z-slab mode (nsol_lsmr_plan_slab): halo planes filled by the caller's neighbour exchange
    bool slab = false;
    int grad_lo = 0, grad_hi = 0, ghost = 0;
    void *halo_v_lo = nullptr, *halo_v_hi = nullptr;   // ghost planes of vhat (blur; the first hi plane also feeds grad)
    void *halo_u_lo = nullptr, *halo_u_hi = nullptr;   // ghost planes of u block 0 (A^T = blur)
    void *halo_uz_lo = nullptr;                        // last plane of the lower neighbour's u_z block (D_z^T)
    void *halo_x_hi = nullptr;                         // first plane of the upper neighbour's x (ADMM: grad x)
    double *ssbuf = nullptr;                           // [1] local / all-reduced sum of squares
    // cooperative single-launch path
    void *taps_dev = nullptr;      // blur taps of all axes in the plan dtype
    int tap_off[3] = {0, 0, 0};
    double *coop_part = nullptr;   // [3][coop_blocks]
    int coop_blocks = 0;           // 0: not initialised, -1: unavailable
    double *coopv_part = nullptr;  // persistent vector solve (lsmr_coopv.cuh): [3][coopv_blocks]
    int coopv_blocks = 0;
};

extern "C" void nsol_lsmr_plan_destroy(nsol_lsmr_plan *pl) {
    if (!pl) return;
    if (pl->ctx) nsol_bind_device(pl->ctx);
    void *ptrs[] = {pl->u, pl->v, pl->h, pl->hbar, pl->x, pl->opbuf, pl->optmp, pl->breg, pl->admm_v, pl->admm_w, pl->admm_c,
                    pl->bbuf, pl->xbuf, pl->stage, pl->part, pl->S, pl->taps_dev, pl->coop_part, pl->coopv_part,
                    pl->halo_v_lo, pl->halo_v_hi, pl->halo_u_lo, pl->halo_u_hi, pl->halo_uz_lo, pl->halo_x_hi, pl->ssbuf};
    for (void *p : ptrs) nsol_plan_free(pl->ctx, p);
    if (pl->own_stream) cudaStreamDestroy(pl->own_stream);
    delete pl;
}

extern "C" int nsol_lsmr_plan_create(nsol_ctx *ctx, const nsol_lsq_desc *desc, nsol_lsmr_plan **out) {
    if (!ctx) return NSOL_EINVAL;
    if (!desc || !out) return nsol_fail(ctx, NSOL_EINVAL, "nsol_lsmr_plan_create: NULL argument");
    *out = nullptr;
    NSOL_CHECK(nsol_bind_device(ctx));
    GridView gv;
    NSOL_CHECK(nsol_grid_view(ctx, &desc->grid, &gv));
    if (gv.batch != 1) return nsol_fail(ctx, NSOL_EINVAL, "lsmr: grid.batch must be 1");
    if (desc->a_op != NSOL_A_BLUR && desc->a_op != NSOL_A_IDENTITY) return nsol_fail(ctx, NSOL_EINVAL, "lsmr: unknown a_op %d", desc->a_op);
    if (desc->b_op < NSOL_B_GRAD || desc->b_op > NSOL_B_NONE) return nsol_fail(ctx, NSOL_EINVAL, "lsmr: unknown b_op %d", desc->b_op);
    nsol_lsmr_plan *pl = new nsol_lsmr_plan();
    pl->ctx = ctx;
    pl->desc = *desc;
    pl->grid = desc->grid;
    pl->gv = gv;
    pl->esz = nsol_dtype_size(gv.dtype);
    if (desc->a_op == NSOL_A_BLUR) {
        for (int a = 0; a < gv.dim; ++a) {
            if (!desc->taps[a] || desc->radius[a] < 0 || desc->radius[a] > 64) {
                delete pl;
                return nsol_fail(ctx, NSOL_EINVAL, "lsmr: blur taps[%d] missing or radius out of range", a);
            }
            pl->taps[a].assign(desc->taps[a], desc->taps[a] + 2 * desc->radius[a] + 1);
        }
    }
    for (int a = 0; a < 3; ++a) pl->desc.taps[a] = nullptr;
    pl->rows_b = desc->b_op == NSOL_B_GRAD ? gv.dim : (desc->b_op == NSOL_B_IDENTITY ? 1 : 0);
    long long want = (gv.n + LSMR_THREADS - 1) / LSMR_THREADS;
    long long cap = ctx->lsmr_blocks > 0 ? ctx->lsmr_blocks : (long long)ctx->sm_count * 8;
    pl->nblocks = (int)(want < cap ? (want > 0 ? want : 1) : cap);
    pl->row_blocks = ((gv.nx + FAST_TH - 1) / FAST_TH) * gv.ny * gv.nz;
    const size_t nb = (size_t)gv.n * pl->esz;
    auto alloc = [&](void **ptr, size_t bytes) -> bool {
        cudaError_t e = nsol_plan_alloc(ctx, ptr, bytes ? bytes : 8);
        if (e != cudaSuccess) {
            nsol_fail(ctx, NSOL_ENOMEM, "lsmr plan: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
            return false;
        }
        pl->bytes += bytes;
        return true;
    };
    bool ok = alloc(&pl->u, nb * (1 + pl->rows_b)) && alloc(&pl->v, nb) && alloc(&pl->h, nb) && alloc(&pl->hbar, nb) && alloc(&pl->x, nb) &&
              alloc(&pl->opbuf, nb) && alloc(&pl->optmp, nb) && alloc(&pl->breg, nb * (pl->rows_b ? pl->rows_b : 1)) &&
              alloc(&pl->admm_v, nb * gv.dim) && alloc(&pl->admm_w, nb * gv.dim) && alloc(&pl->bbuf, nb) && alloc(&pl->xbuf, nb) &&
              alloc((void **)&pl->part, sizeof(double) * (size_t)(pl->nblocks > pl->row_blocks ? pl->nblocks : pl->row_blocks) * 2) && alloc((void **)&pl->S, sizeof(LsmrScalars));
    if (!ok) {
        nsol_lsmr_plan_destroy(pl);
        return NSOL_ENOMEM;
    }
    *out = pl;
    return NSOL_OK;
}

extern "C" size_t nsol_lsmr_plan_bytes(const nsol_lsmr_plan *pl) { return pl ? pl->bytes + pl->stage_bytes : 0; }

template <typename T>
static LsqGeom<T> make_geom(const nsol_lsmr_plan *pl) {
    const GridView &gv = pl->gv;
    LsqGeom<T> g;
    g.n = gv.n;
    g.nx = gv.nx;
    g.ny = gv.ny;
    g.nz = gv.nz;
    g.dim = gv.dim;
    g.b_op = pl->desc.b_op;
    g.slab = pl->slab ? 1 : 0;
    g.grad_lo = pl->grad_lo;
    g.grad_hi = pl->grad_hi;
    g.ghost = pl->ghost;
    for (int k = 0; k < 3; ++k) {
        g.stride[k] = 0;
        g.extent[k] = 1;
        g.axis[k] = 0;
        g.w[k] = T(0);
    }
    for (int k = 0; k < gv.dim; ++k) {
        const int ax = (k == 0) ? 0 : ((gv.dim == 3 && k == 1) ? 1 : 2);
        g.axis[k] = ax;
        g.stride[k] = ax == 0 ? 1 : (ax == 1 ? gv.nx : (long long)gv.nx * gv.ny);
        g.extent[k] = ax == 0 ? gv.nx : (ax == 1 ? gv.ny : gv.nz);
        g.w[k] = (T)gv.w[k];
    }
    return g;
}

// taps of numpy axis `ax` as a kernel argument (r = -1: identity, no blur)
template <typename T>
static TapArgs<T> lsq_taps(const nsol_lsmr_plan *pl, int ax) {
    TapArgs<T> t;
    t.r = -1;
    if (pl->desc.a_op == NSOL_A_BLUR && ax >= 0) {
        t.r = pl->desc.radius[ax];
        for (int k = 0; k <= 2 * t.r; ++k) t.t[k] = (T)pl->taps[ax][k];
    }
    return t;
}

#include "lsmr_fastv.cuh"
#include "lsmr_fused3d.cuh"
#include "lsmr_fused2d_v2.cuh"

// blur passes along every numpy axis except the last (x): in -> optmp (-> opbuf); *result is what the
// consumer's fused x-pass reads (in itself for 1-D problems or A = identity)
template <typename T>
static int lsq_blur_front(nsol_lsmr_plan *pl, const LsqGeom<T> &g, const void *in, const void **result, cudaStream_t s,
                          const void *halo_lo = nullptr, const void *halo_hi = nullptr) {
    *result = in;
    if (pl->desc.a_op != NSOL_A_BLUR) return NSOL_OK;
    const dim3 grid((g.nx + FAST_TH - 1) / FAST_TH, g.ny, g.nz);
    const void *src = in;
    for (int ax = 0; ax + 1 < g.dim; ++ax) {
        void *dst = (src == pl->optmp) ? pl->opbuf : pl->optmp;
        const int kaxis = (g.dim == 3 && ax == 1) ? 1 : 2;
        if (fastv_ok(pl)) {
            constexpr int VEC = FastvCfg<T>::VEC;
            dim3 vgrid((g.nx / VEC + FAST_TH - 1) / FAST_TH, g.ny, g.nz);
            if (kaxis == 1) vgrid.y = (g.ny + FASTV_ROWS - 1) / FASTV_ROWS;
            else vgrid.z = (g.nz + FASTV_ROWS - 1) / FASTV_ROWS;
            const FastvGeom<T> fg = make_fastv_geom<T>(g);
            FASTV_SWITCH_R(lsq_radius(pl, ax), (fastv_blur_pass_kernel<T, R, VEC><<<vgrid, FAST_TH, 0, s>>>(
                                                   fg, lsq_taps_r<T, R>(pl, ax), kaxis, (const T *)src, (T *)dst, (const T *)halo_lo,
                                                   (const T *)halo_hi)));
        } else {
            fast_blur_pass_kernel<T><<<grid, FAST_TH, 0, s>>>(g, lsq_taps<T>(pl, ax), kaxis, (const T *)src, (T *)dst, (const T *)halo_lo,
                                                              (const T *)halo_hi);
        }
        NSOL_LAUNCH_CHECK(pl->ctx);
        src = dst;
    }
    *result = src;
    return NSOL_OK;
}

// forward / adjoint / update phases: vector kernels when they apply, the generic ones otherwise.
// *nparts = number of partial sums written to pl->part.
template <typename T>
static int lsq_launch_fwd(nsol_lsmr_plan *pl, const LsqGeom<T> &g, const void *halo_lo, const void *halo_hi, const void *v_hi,
                          cudaStream_t s, int *nparts) {
    T *u = (T *)pl->u;
    const T *v = (const T *)pl->v;
    const int ax = g.dim - 1;
    if (fused2d_ok(pl, g.b_op)) return pl->ctx->lsmr_fuse2d == 3 ? fused2d_launch<T>(pl, true, 0, s, nparts) : fused2d_v2_launch<T>(pl, true, 0, s, nparts);
    if (fused3d_ok(pl, g.b_op)) return fused3d_launch<T>(pl, true, 0, s, nparts);
    const void *op = nullptr;
    NSOL_CHECK(lsq_blur_front<T>(pl, g, v, &op, s, halo_lo, halo_hi));
    if (fastv_ok(pl)) {
        constexpr int VEC = FastvCfg<T>::VEC;
        const dim3 vgrid((g.nx / VEC + FAST_TH - 1) / FAST_TH, g.ny, g.nz);
        const FastvGeom<T> fg = make_fastv_geom<T>(g);
        const int r = lsq_radius(pl, ax) < 0 ? 0 : lsq_radius(pl, ax);
        FASTV_SWITCH_R(r, (fastv_fwd_kernel<T, R, VEC><<<vgrid, FAST_TH, 0, s>>>(fg, pl->S, lsq_taps_r<T, R>(pl, lsq_radius(pl, ax) < 0 ? -1 : ax),
                                                                                  (const T *)op, v, u, pl->part, (const T *)v_hi)));
        *nparts = (int)(vgrid.x * vgrid.y * vgrid.z);
    } else {
        const dim3 rgrid((g.nx + FAST_TH - 1) / FAST_TH, g.ny, g.nz);
        fast_fwd_kernel<T><<<rgrid, FAST_TH, 0, s>>>(g, pl->S, lsq_taps<T>(pl, ax), (const T *)op, v, u, pl->part, (const T *)v_hi);
        *nparts = pl->row_blocks;
    }
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

template <typename T>
static int lsq_launch_adj(nsol_lsmr_plan *pl, const LsqGeom<T> &g, const void *halo_lo, const void *halo_hi, int first, const void *uz_lo,
                          cudaStream_t s, int *nparts) {
    const T *u = (const T *)pl->u;
    T *v = (T *)pl->v;
    const int ax = g.dim - 1;
    if (fused2d_ok(pl, g.b_op)) return pl->ctx->lsmr_fuse2d == 3 ? fused2d_launch<T>(pl, false, first, s, nparts) : fused2d_v2_launch<T>(pl, false, first, s, nparts);
    if (fused3d_ok(pl, g.b_op)) return fused3d_launch<T>(pl, false, first, s, nparts);
    const void *op = nullptr;
    NSOL_CHECK(lsq_blur_front<T>(pl, g, u, &op, s, halo_lo, halo_hi));       // A^T = A (same mask, periodic)
    if (fastv_ok(pl)) {
        constexpr int VEC = FastvCfg<T>::VEC;
        const dim3 vgrid((g.nx / VEC + FAST_TH - 1) / FAST_TH, g.ny, g.nz);
        const FastvGeom<T> fg = make_fastv_geom<T>(g);
        const int r = lsq_radius(pl, ax) < 0 ? 0 : lsq_radius(pl, ax);
        FASTV_SWITCH_R(r, (fastv_adj_kernel<T, R, VEC><<<vgrid, FAST_TH, 0, s>>>(fg, pl->S, lsq_taps_r<T, R>(pl, lsq_radius(pl, ax) < 0 ? -1 : ax),
                                                                                  (const T *)op, u, v, pl->part, first, (const T *)uz_lo)));
        *nparts = (int)(vgrid.x * vgrid.y * vgrid.z);
    } else {
        const dim3 rgrid((g.nx + FAST_TH - 1) / FAST_TH, g.ny, g.nz);
        fast_adj_kernel<T><<<rgrid, FAST_TH, 0, s>>>(g, pl->S, lsq_taps<T>(pl, ax), (const T *)op, u, v, pl->part, first, (const T *)uz_lo);
        *nparts = pl->row_blocks;
    }
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

template <typename T>
static int lsq_launch_update(nsol_lsmr_plan *pl, const LsqGeom<T> &g, cudaStream_t s, int *nparts) {
    T *v = (T *)pl->v, *h = (T *)pl->h, *hbar = (T *)pl->hbar, *x = (T *)pl->x;
    constexpr int VEC = FastvCfg<T>::VEC;
    if (fastv_ok(pl) && g.n % VEC == 0) {
        const long long nvec = g.n / VEC;
        long long want = (nvec + 2 * LSMR_THREADS - 1) / (2 * LSMR_THREADS);
        const long long cap = (long long)pl->ctx->sm_count * 8;
        const int nb = (int)(want < cap ? (want > 0 ? want : 1) : cap);
        fastv_update_kernel<T, VEC><<<nb, LSMR_THREADS, 0, s>>>(nvec, pl->S, v, h, hbar, x, pl->part);
        *nparts = nb;
    } else {
        lsmr_update_kernel<T><<<pl->nblocks, LSMR_THREADS, 0, s>>>(g.n, pl->S, v, h, hbar, x, pl->part);
        *nparts = pl->nblocks;
    }
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

// number of blocks of a flat streaming kernel over nvec 16-byte vectors
static int lsq_flat_blocks(const nsol_lsmr_plan *pl, long long nvec) {
    long long want = (nvec + LSMR_THREADS - 1) / LSMR_THREADS;
    const long long cap = pl->ctx->lsmr_blocks > 0 ? pl->ctx->lsmr_blocks : (long long)pl->ctx->sm_count * 8;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <typename T>
static int lsq_launch_rhs(nsol_lsmr_plan *pl, const LsqGeom<T> &g, int rows_b, const void *b, const void *breg, double sa, cudaStream_t s, int *nparts,
                          const double *sa_dev = nullptr) {
    constexpr int VEC = FastvCfg<T>::VEC;
    if (fastv_ok(pl)) {
        const long long nvec = g.n / VEC;
        const int nb = lsq_flat_blocks(pl, nvec * (1 + rows_b));
        fastv_rhs_kernel<T, VEC><<<nb, LSMR_THREADS, 0, s>>>(nvec, rows_b, (const T *)b, (const T *)breg, (T)sa, (T *)pl->u, pl->part, sa_dev);
        *nparts = nb;
    } else {
        lsmr_rhs_kernel<T><<<pl->nblocks, LSMR_THREADS, 0, s>>>(g, rows_b, (const T *)b, (const T *)breg, sa, (T *)pl->u, pl->part, sa_dev);
        *nparts = pl->nblocks;
    }
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

template <typename T>
static int lsq_launch_init_vectors(nsol_lsmr_plan *pl, const LsqGeom<T> &g, cudaStream_t s) {
    constexpr int VEC = FastvCfg<T>::VEC;
    T *v = (T *)pl->v, *h = (T *)pl->h, *hbar = (T *)pl->hbar, *x = (T *)pl->x;
    if (fastv_ok(pl)) fastv_init_vectors_kernel<T, VEC><<<lsq_flat_blocks(pl, g.n / VEC), LSMR_THREADS, 0, s>>>(g.n / VEC, pl->S, v, h, hbar, x);
    else lsmr_init_vectors_kernel<T><<<pl->nblocks, LSMR_THREADS, 0, s>>>(g.n, pl->S, v, h, hbar, x);
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

template <typename T>
static int lsq_launch_clip(nsol_lsmr_plan *pl, const LsqGeom<T> &g, void *out, double lo, double hi, cudaStream_t s) {
    constexpr int VEC = FastvCfg<T>::VEC;
    if (fastv_ok(pl)) fastv_clip_kernel<T, VEC><<<lsq_flat_blocks(pl, g.n / VEC), LSMR_THREADS, 0, s>>>(g.n / VEC, (const T *)pl->x, (T *)out, lo, hi);
    else clip_kernel<T><<<pl->nblocks, LSMR_THREADS, 0, s>>>(g.n, (const T *)pl->x, (T *)out, lo, hi);
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

template <typename T>
static int lsq_launch_shrink(nsol_lsmr_plan *pl, const LsqGeom<T> &g, const void *x, const void *w_in, double ell, void *v, void *w, void *breg,
                             const void *x_hi, int plain, cudaStream_t s) {
    constexpr int VEC = FastvCfg<T>::VEC;
    if (fastv_ok(pl)) {
        const dim3 vgrid((g.nx / VEC + FAST_TH - 1) / FAST_TH, g.ny, g.nz);
        fastv_shrink_kernel<T, VEC><<<vgrid, FAST_TH, 0, s>>>(make_fastv_geom<T>(g), (const T *)x, (const T *)w_in, (T)ell, (T *)v, (T *)w, (T *)breg,
                                                              (const T *)x_hi, plain, (const T *)pl->admm_c);
    } else {
        admm_shrink_kernel<T><<<pl->nblocks, LSMR_THREADS, 0, s>>>(g, (const T *)x, (const T *)w_in, (T)ell, (T *)v, (T *)w, (T *)breg, (const T *)x_hi, plain,
                                                                   (const T *)pl->admm_c);
    }
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

template <typename T>
static int lsmr_solve_t(nsol_lsmr_plan *pl, double alpha, const void *b_dev, const void *breg_dev, int maxiter, double lo, double hi,
                        void *x_out, cudaStream_t s, const double *sa_dev = nullptr) {
    nsol_ctx *ctx = pl->ctx;
    const LsqGeom<T> g = make_geom<T>(pl);
    // alpha <= EPS: plain system A x = b (tikhonov_linear_solver.py:241-248)
    const bool use_b = alpha > 1e-10 && pl->rows_b > 0;
    LsqGeom<T> ge = g;
    if (!use_b) ge.b_op = NSOL_B_NONE;
    const int rows_b = use_b ? pl->rows_b : 0;
    const double sa = use_b ? sqrt(alpha) : 0.0;
    double *part = pl->part;

    int rhs_parts = 0;
    NSOL_CHECK(lsq_launch_rhs<T>(pl, ge, rows_b, b_dev, breg_dev, sa, s, &rhs_parts, sa_dev));
    lsmr_scalar_init_beta<<<1, 1024, 0, s>>>(pl->S, part, rhs_parts, sa, maxiter, sa_dev);
    NSOL_LAUNCH_CHECK(ctx);
    if (g.ny > 65535 || g.nz > 65535)
        return nsol_fail(ctx, NSOL_EINVAL, "lsmr (multi-kernel path): more than 65535 rows along y or z are not supported");
    int rparts = 0;
    NSOL_CHECK(lsq_launch_adj<T>(pl, ge, nullptr, nullptr, 1, nullptr, s, &rparts));
    lsmr_scalar_init_alpha<<<1, 1024, 0, s>>>(pl->S, part, rparts);
    NSOL_LAUNCH_CHECK(ctx);
    NSOL_CHECK(lsq_launch_init_vectors<T>(pl, ge, s));
    for (int it = 0; it < maxiter; ++it) {
        NSOL_CHECK(lsq_launch_fwd<T>(pl, ge, nullptr, nullptr, nullptr, s, &rparts));
        lsmr_scalar_beta<<<1, 1024, 0, s>>>(pl->S, part, rparts);
        NSOL_LAUNCH_CHECK(ctx);
        NSOL_CHECK(lsq_launch_adj<T>(pl, ge, nullptr, nullptr, 0, nullptr, s, &rparts));
        lsmr_scalar_alpha<<<1, 1024, 0, s>>>(pl->S, part, rparts);
        NSOL_LAUNCH_CHECK(ctx);
        NSOL_CHECK(lsq_launch_update<T>(pl, ge, s, &rparts));
        lsmr_scalar_tests<<<1, 1024, 0, s>>>(pl->S, part, rparts);
        NSOL_LAUNCH_CHECK(ctx);
    }
    NSOL_CHECK(lsq_launch_clip<T>(pl, ge, x_out, lo, hi, s));
    return NSOL_OK;
}

#include "lsmr_coop.cuh"
#include "lsmr_coopv.cuh"

// one-time setup of the cooperative path: taps on the device, co-resident grid size, partial sums
template <typename T>
static int coop_prepare(nsol_lsmr_plan *pl) {
    nsol_ctx *ctx = pl->ctx;
    if (pl->coop_blocks != 0) return NSOL_OK;
    int coop = 0;
    NSOL_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!coop) {
        pl->coop_blocks = -1;
        return NSOL_OK;
    }
    std::vector<T> taps;
    for (int a = 0; a < pl->gv.dim; ++a) {
        pl->tap_off[a] = (int)taps.size();
        for (double t : pl->taps[a]) taps.push_back((T)t);
    }
    if (taps.empty()) taps.push_back(T(0));
    NSOL_CUDA(ctx, cudaMalloc(&pl->taps_dev, taps.size() * sizeof(T)));
    NSOL_CUDA(ctx, cudaMemcpy(pl->taps_dev, taps.data(), taps.size() * sizeof(T), cudaMemcpyHostToDevice));
    int per_sm = 0;
    NSOL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lsmr_coop_kernel<T>, LSMR_THREADS, 0));
    if (per_sm < 1) {
        pl->coop_blocks = -1;
        return NSOL_OK;
    }
    if (per_sm > 4) per_sm = 4;   // grid.sync() cost grows with the number of blocks
    long long want = (pl->gv.n + LSMR_THREADS - 1) / LSMR_THREADS;
    long long cap = (long long)per_sm * ctx->sm_count;
    if (ctx->lsmr_blocks > 0 && ctx->lsmr_blocks < cap) cap = ctx->lsmr_blocks;
    const int blocks = (int)(want < cap ? (want > 0 ? want : 1) : cap);
    NSOL_CUDA(ctx, cudaMalloc((void **)&pl->coop_part, sizeof(double) * 3 * (size_t)blocks));
    pl->coop_blocks = blocks;
    return NSOL_OK;
}

// whole solve (admm_iters == 0) or whole ADMM run (admm_iters > 0) in one cooperative launch
template <typename T>
static int lsmr_solve_coop(nsol_lsmr_plan *pl, double alpha, const void *b_dev, void *breg_dev, int maxiter, double lo, double hi,
                           void *x_out, int admm_iters, double ell, cudaStream_t s) {
    nsol_ctx *ctx = pl->ctx;
    CoopArgs<T> a;
    a.g = make_geom<T>(pl);
    const bool use_b = alpha > 1e-10 && pl->rows_b > 0;
    if (!use_b) a.g.b_op = NSOL_B_NONE;
    a.rows_b = use_b ? pl->rows_b : 0;
    a.a_blur = pl->desc.a_op == NSOL_A_BLUR;
    long long acc = 1;
    for (int ax = 2; ax >= 0; --ax) {
        a.np_stride[ax] = 0;
        a.np_extent[ax] = 1;
        a.radius[ax] = 0;
        a.tap_off[ax] = 0;
    }
    for (int ax = pl->gv.dim - 1; ax >= 0; --ax) {
        a.np_stride[ax] = acc;
        a.np_extent[ax] = (int)pl->grid.shape[ax];
        acc *= pl->grid.shape[ax];
        a.radius[ax] = a.a_blur ? pl->desc.radius[ax] : 0;
        a.tap_off[ax] = pl->tap_off[ax];
    }
    a.taps = (const T *)pl->taps_dev;
    a.b = (const T *)b_dev;
    a.breg = use_b ? (T *)breg_dev : nullptr;
    a.u = (T *)pl->u; a.v = (T *)pl->v; a.h = (T *)pl->h; a.hbar = (T *)pl->hbar; a.x = (T *)pl->x;
    a.opbuf = (T *)pl->opbuf; a.optmp = (T *)pl->optmp;
    a.xout = (T *)x_out;
    a.part = pl->coop_part;
    a.S = pl->S;
    a.sqrt_alpha = use_b ? sqrt(alpha) : 0.0;
    a.lo = lo; a.hi = hi;
    a.maxiter = maxiter;
    a.admm_iters = admm_iters;
    a.ell = (T)ell;
    a.admm_v = (T *)pl->admm_v; a.admm_w = (T *)pl->admm_w;
    a.admm_c = (const T *)pl->admm_c;
    void *params[] = {(void *)&a};
    NSOL_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)lsmr_coop_kernel<T>, dim3(pl->coop_blocks), dim3(LSMR_THREADS), params, 0, s));
    ctx->launches++;
    return NSOL_OK;
}

// Path choice (measured on B200, profiles/r1_lsmr_paths.md): the single cooperative launch wins only for
// very small vectors (sync floor ~23 us per inner iteration at 256^2); from 512^2 on, one row-mapped
// kernel per phase + CUDA-graph replay is faster (40 vs 44 us) because more threads are in flight than a
// co-resident grid allows.
static int lsmr_use_coop(nsol_lsmr_plan *pl) {
    if (pl->slab) return nsol_fail(pl->ctx, NSOL_ESTATE, "lsmr: a z-slab plan is driven phase by phase (nsol_lsmr_slab_phase)");
    if (pl->ctx->lsmr_path == 1 || pl->ctx->lsmr_path == 3) return 0;
    if (pl->ctx->lsmr_path == 0 && pl->gv.n > (1ll << 17)) return 0;
    int rc = pl->gv.dtype == NSOL_F32 ? coop_prepare<float>(pl) : coop_prepare<double>(pl);
    if (rc != NSOL_OK) return rc;
    return pl->coop_blocks > 0 ? 1 : 0;
}

static int lsmr_solve_any(nsol_lsmr_plan *pl, double alpha, const void *b_dev, const void *breg_dev, int maxiter, double lo, double hi,
                          void *x_out, cudaStream_t s) {
    if (coopv_ok(pl)) {      // small / mid-size problem: the whole solve as one persistent vector launch
        const int rc = pl->gv.dtype == NSOL_F32 ? lsmr_solve_coopv<float>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s)
                                                : lsmr_solve_coopv<double>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s);
        if (rc != NSOL_ESTATE) return rc;
    }
    const int coop = lsmr_use_coop(pl);
    if (coop < 0) return coop;
    if (coop) {
        if (pl->gv.dtype == NSOL_F32) return lsmr_solve_coop<float>(pl, alpha, b_dev, (void *)breg_dev, maxiter, lo, hi, x_out, 0, 0.0, s);
        return lsmr_solve_coop<double>(pl, alpha, b_dev, (void *)breg_dev, maxiter, lo, hi, x_out, 0, 0.0, s);
    }
    if (pl->gv.dtype == NSOL_F32) return lsmr_solve_t<float>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s);
    return lsmr_solve_t<double>(pl, alpha, b_dev, breg_dev, maxiter, lo, hi, x_out, s);
}

extern "C" int nsol_lsmr_solve_dev(nsol_lsmr_plan *pl, double alpha, const void *b_dev, const void *b_reg_dev, int maxiter, double lo,
                                   double hi, void *x_dev, int *itn_out, int *istop_out, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!b_dev || !x_dev) return nsol_fail(ctx, NSOL_EINVAL, "lsmr solve: NULL array");
    if (maxiter < 0) return nsol_fail(ctx, NSOL_EINVAL, "lsmr solve: maxiter must be >= 0");
    if (!(alpha >= 0.0)) return nsol_fail(ctx, NSOL_EINVAL, "lsmr solve: alpha must be >= 0");
    NSOL_CHECK(nsol_bind_device(ctx));
    cudaStream_t st = (cudaStream_t)s;
    NSOL_CHECK(lsmr_solve_any(pl, alpha, b_dev, b_reg_dev, maxiter, lo, hi, x_dev, st));
    if (itn_out || istop_out) {
        LsmrScalars hS;
        NSOL_CUDA(ctx, cudaMemcpyAsync(&hS, pl->S, sizeof(hS), cudaMemcpyDeviceToHost, st));
        NSOL_CUDA(ctx, cudaStreamSynchronize(st));
        if (itn_out) *itn_out = hS.itn;
        if (istop_out) *istop_out = hS.istop;
    }
    return NSOL_OK;
}

static int lsq_ensure_stage(nsol_lsmr_plan *pl, size_t bytes) {
    if (pl->stage_bytes >= bytes) return NSOL_OK;
    nsol_ctx *ctx = pl->ctx;
    if (pl->stage) {
        NSOL_CUDA(ctx, cudaDeviceSynchronize());
        NSOL_CUDA(ctx, cudaFree(pl->stage));
        pl->stage = nullptr;
        pl->stage_bytes = 0;
    }
    NSOL_CUDA(ctx, cudaMalloc(&pl->stage, bytes));
    pl->stage_bytes = bytes;
    return NSOL_OK;
}

static int lsq_download(nsol_lsmr_plan *pl, const void *x_dev, double x_scale, double *x_host, cudaStream_t st) {
    nsol_ctx *ctx = pl->ctx;
    const size_t n = (size_t)pl->gv.n;
    NSOL_CHECK(lsq_ensure_stage(pl, n * sizeof(double)));
    NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)n, pl->gv.dtype, x_dev, NSOL_F64, pl->stage, x_scale, 0, st));
    NSOL_CUDA(ctx, cudaMemcpyAsync(x_host, pl->stage, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    NSOL_CUDA(ctx, cudaStreamSynchronize(st));
    return NSOL_OK;
}

// TikhonovLinearSolver.run() + get_x(): b, b_reg unscaled; solver works on b/x_scale, b_reg/x_scale
// (linear_solver.py:80-89, tikhonov_linear_solver.py:88-89); result multiplied back (solver.py:117-118).
extern "C" int nsol_tikhonov_run_host(nsol_lsmr_plan *pl, double alpha, double in_scale, double out_scale, const double *b_host,
                                      const double *b_reg_host, int maxiter, double lo, double hi, double *x_host, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!b_host || !x_host) return nsol_fail(ctx, NSOL_EINVAL, "tikhonov run: NULL array");
    if (in_scale == 0.0 || out_scale == 0.0) return nsol_fail(ctx, NSOL_EINVAL, "tikhonov run: scales must be non-zero");
    NSOL_CHECK(nsol_bind_device(ctx));
    cudaStream_t st = (cudaStream_t)s;
    const size_t n = (size_t)pl->gv.n;
    const size_t nreg = b_reg_host ? n * (pl->rows_b ? pl->rows_b : 1) : 0;
    NSOL_CHECK(lsq_ensure_stage(pl, (n + nreg) * sizeof(double)));
    double *sb = (double *)pl->stage;
    NSOL_CUDA(ctx, cudaMemcpyAsync(sb, b_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)n, NSOL_F64, sb, pl->gv.dtype, pl->bbuf, in_scale, 1, st));
    const void *breg = nullptr;
    if (b_reg_host && pl->rows_b) {
        NSOL_CUDA(ctx, cudaMemcpyAsync(sb + n, b_reg_host, nreg * sizeof(double), cudaMemcpyHostToDevice, st));
        NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)nreg, NSOL_F64, sb + n, pl->gv.dtype, pl->breg, in_scale, 1, st));
        breg = pl->breg;
    }
    NSOL_CHECK(lsmr_solve_any(pl, alpha, pl->bbuf, breg, maxiter, lo, hi, pl->xbuf, st));
    return lsq_download(pl, pl->xbuf, out_scale, x_host, st);
}

template <typename T>
static int admm_iterations_t(nsol_lsmr_plan *pl, double alpha, double rho, int iterations, int iter_max, const void *b_dev, void *x_dev,
                             double x_scale, double *iterates_host, cudaStream_t s) {
    nsol_ctx *ctx = pl->ctx;
    const LsqGeom<T> g = make_geom<T>(pl);
    const size_t n = (size_t)pl->gv.n;
    T *v = (T *)pl->admm_v, *w = (T *)pl->admm_w, *breg = (T *)pl->breg;
    // v = B(x0) - b_reg (b_reg = 0), w = 0  (admm_linear_solver.py:171-172): shrink with ell < 0 keeps v = t
    // -> do it explicitly: v = grad(x0), w = 0, b_reg_next = v - w = v
    if (pl->admm_c) {
        if (g.ny > 65535 || g.nz > 65535) return nsol_fail(ctx, NSOL_EINVAL, "admm: more than 65535 rows along y or z are not supported with b_reg");
        NSOL_CHECK(lsq_launch_shrink<T>(pl, g, x_dev, nullptr, 0.0, v, w, breg, nullptr, 1, s));    // v = B x0 - c, w = 0, b_reg = v + c
    } else {
        NSOL_CHECK(nsol_grad(ctx, &pl->grid, x_dev, v, s));
        NSOL_CUDA(ctx, cudaMemsetAsync(w, 0, n * pl->gv.dim * sizeof(T), s));
        NSOL_CUDA(ctx, cudaMemcpyAsync(breg, v, n * pl->gv.dim * sizeof(T), cudaMemcpyDeviceToDevice, s));
    }
    if (iterates_host) NSOL_CHECK(lsq_download(pl, x_dev, x_scale, iterates_host, s));
    auto outer_iteration = [&](cudaStream_t st) -> int {
        // x <- clip(lsmr([A; sqrt(rho) B], [b; sqrt(rho)(v - w)]), 0, inf)   (:205, :220-237; x0 is not passed to lsmr)
        NSOL_CHECK(lsmr_solve_any(pl, rho, b_dev, breg, iter_max, 0.0, INFINITY, x_dev, st));
        // t = B x + w ; v = prox_g(t, alpha/rho) ; w = t - v   (:208-216)
        NSOL_CHECK(lsq_launch_shrink<T>(pl, g, x_dev, w, alpha / rho, v, w, breg, nullptr, 0, st));
        return NSOL_OK;
    };
    if (coopv_ok(pl)) {
        // persistent vector solve: two launches per outer iteration (solve, shrink) -- nothing left for a graph to save
        for (int it = 0; it < iterations; ++it) {
            NSOL_CHECK(outer_iteration(s));
            if (iterates_host) NSOL_CHECK(lsq_download(pl, x_dev, x_scale, iterates_host + (size_t)(it + 1) * n, s));
        }
        return NSOL_OK;
    }
    const int coop = lsmr_use_coop(pl);
    if (coop < 0) return coop;
    if (coop && iterations > 0) {
        // cooperative path: the whole ADMM run is ONE launch (or one per outer iteration when the
        // observer wants every iterate)
        if (!iterates_host) return lsmr_solve_coop<T>(pl, rho, b_dev, breg, iter_max, 0.0, INFINITY, x_dev, iterations, alpha / rho, s);
        for (int it = 0; it < iterations; ++it) {
            NSOL_CHECK(lsmr_solve_coop<T>(pl, rho, b_dev, breg, iter_max, 0.0, INFINITY, x_dev, 1, alpha / rho, s));
            NSOL_CHECK(lsq_download(pl, x_dev, x_scale, iterates_host + (size_t)(it + 1) * n, s));
        }
        return NSOL_OK;
    }
    if (iterates_host || iterations < 2) {
        // observer attached (one download per outer iteration), or nothing to replay
        for (int it = 0; it < iterations; ++it) {
            NSOL_CHECK(outer_iteration(s));
            if (iterates_host) NSOL_CHECK(lsq_download(pl, x_dev, x_scale, iterates_host + (size_t)(it + 1) * n, s));
        }
        return NSOL_OK;
    }
    // The launch sequence of one outer iteration is fixed (LSMR stops early through device-side
    // flags, not by skipping launches), so it is captured once into a CUDA graph and replayed:
    // ~10*iter_max small kernels per outer iteration stop paying individual launch latency.
    const int64_t l0 = ctx->launches;
    cudaGraph_t graph = nullptr;
    NSOL_CUDA(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc = outer_iteration(s);
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (rc != NSOL_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (ce != cudaSuccess) return nsol_fail(ctx, NSOL_ECUDA, "admm: graph capture failed: %s", cudaGetErrorString(ce));
    const int64_t per_iteration = ctx->launches - l0;
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return nsol_fail(ctx, NSOL_ECUDA, "admm: graph instantiation failed: %s", cudaGetErrorString(ce));
    for (int it = 0; it < iterations && ce == cudaSuccess; ++it) ce = cudaGraphLaunch(exec, s);
    ctx->launches += per_iteration * (iterations - 1);
    // the executable graph must outlive its launches
    cudaError_t se = cudaStreamSynchronize(s);
    cudaGraphExecDestroy(exec);
    if (ce != cudaSuccess) return nsol_fail(ctx, NSOL_ECUDA, "admm: graph launch failed: %s", cudaGetErrorString(ce));
    if (se != cudaSuccess) return nsol_fail(ctx, NSOL_ECUDA, "admm: %s", cudaGetErrorString(se));
    return NSOL_OK;
}

// stream the ADMM run is ordered on: the caller's, or a plan-owned blocking stream (which is
// implicitly ordered with the legacy default stream) when the caller passed NULL
static int admm_stream(nsol_lsmr_plan *pl, nsol_stream s, cudaStream_t *out) {
    if (s) {
        *out = (cudaStream_t)s;
        return NSOL_OK;
    }
    if (!pl->own_stream) NSOL_CUDA(pl->ctx, cudaStreamCreate(&pl->own_stream));
    *out = pl->own_stream;
    return NSOL_OK;
}

static int admm_check(nsol_lsmr_plan *pl, double alpha, double rho, int iterations, int iter_max) {
    nsol_ctx *ctx = pl->ctx;
    if (pl->desc.b_op != NSOL_B_GRAD) return nsol_fail(ctx, NSOL_EINVAL, "admm: the plan's B must be the gradient");
    if (!(rho > 1e-10)) return nsol_fail(ctx, NSOL_EINVAL, "admm: rho must be > 0");
    if (!(alpha >= 0.0)) return nsol_fail(ctx, NSOL_EINVAL, "admm: alpha must be >= 0");
    if (iterations < 0 || iter_max < 0) return nsol_fail(ctx, NSOL_EINVAL, "admm: iterations and iter_max must be >= 0");
    return NSOL_OK;
}

extern "C" int nsol_admm_run_dev(nsol_lsmr_plan *pl, double alpha, double rho, int iterations, int iter_max, const void *b_dev,
                                 const void *x0_dev, void *x_dev, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!b_dev || !x0_dev || !x_dev) return nsol_fail(ctx, NSOL_EINVAL, "admm run: NULL array");
    NSOL_CHECK(admm_check(pl, alpha, rho, iterations, iter_max));
    NSOL_CHECK(nsol_bind_device(ctx));
    cudaStream_t st = nullptr;
    NSOL_CHECK(admm_stream(pl, s, &st));
    if (x_dev != x0_dev) NSOL_CUDA(ctx, cudaMemcpyAsync(x_dev, x0_dev, (size_t)pl->gv.n * pl->esz, cudaMemcpyDeviceToDevice, st));
    if (pl->gv.dtype == NSOL_F32) return admm_iterations_t<float>(pl, alpha, rho, iterations, iter_max, b_dev, x_dev, 1.0, nullptr, st);
    return admm_iterations_t<double>(pl, alpha, rho, iterations, iter_max, b_dev, x_dev, 1.0, nullptr, st);
}

extern "C" int nsol_admm_run_host(nsol_lsmr_plan *pl, double alpha, double rho, int iterations, int iter_max, double in_scale,
                                  double out_scale, const double *b_host, const double *x0_host, double *x_host, double *iterates_host,
                                  nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!b_host || !x0_host || !x_host) return nsol_fail(ctx, NSOL_EINVAL, "admm run: NULL array");
    if (in_scale == 0.0 || out_scale == 0.0) return nsol_fail(ctx, NSOL_EINVAL, "admm run: scales must be non-zero");
    NSOL_CHECK(admm_check(pl, alpha, rho, iterations, iter_max));
    NSOL_CHECK(nsol_bind_device(ctx));
    cudaStream_t st = nullptr;
    NSOL_CHECK(admm_stream(pl, s, &st));
    const size_t n = (size_t)pl->gv.n;
    NSOL_CHECK(lsq_ensure_stage(pl, 2 * n * sizeof(double)));
    double *sb = (double *)pl->stage;
    // b / x_scale, x0 / x_scale  (linear_solver.py:83, solver.py:37)
    NSOL_CUDA(ctx, cudaMemcpyAsync(sb, b_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    NSOL_CUDA(ctx, cudaMemcpyAsync(sb + n, x0_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)n, NSOL_F64, sb, pl->gv.dtype, pl->bbuf, in_scale, 1, st));
    NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)n, NSOL_F64, sb + n, pl->gv.dtype, pl->xbuf, in_scale, 1, st));
    int rc;
    if (pl->gv.dtype == NSOL_F32) rc = admm_iterations_t<float>(pl, alpha, rho, iterations, iter_max, pl->bbuf, pl->xbuf, out_scale, iterates_host, st);
    else rc = admm_iterations_t<double>(pl, alpha, rho, iterations, iter_max, pl->bbuf, pl->xbuf, out_scale, iterates_host, st);
    NSOL_CHECK(rc);
    return lsq_download(pl, pl->xbuf, out_scale, x_host, st);
}

// The ADMM solver's own b_reg (ADMMLinearSolver(..., b_reg=...), nsol/admm_linear_solver.py:100): dim * N float64 host
// values, stored as b_reg / in_scale in the plan's dtype and used by every following ADMM run
// (v0 = B x0 - b_reg, t = B x + w - b_reg, Tikhonov b_reg = v - w + b_reg; :171, :208, :222).  NULL = 0 (default).
extern "C" int nsol_admm_set_b_reg_host(nsol_lsmr_plan *pl, const double *b_reg_host, double in_scale, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    NSOL_CHECK(nsol_bind_device(ctx));
    if (!b_reg_host) {
        if (pl->admm_c) {
            NSOL_CUDA(ctx, cudaDeviceSynchronize());
            cudaFree(pl->admm_c);
            pl->admm_c = nullptr;
        }
        return NSOL_OK;
    }
    if (in_scale == 0.0) return nsol_fail(ctx, NSOL_EINVAL, "admm b_reg: scale must be non-zero");
    cudaStream_t st = nullptr;
    NSOL_CHECK(admm_stream(pl, s, &st));
    const size_t nreg = (size_t)pl->gv.n * pl->gv.dim;
    if (!pl->admm_c) {
        NSOL_CUDA(ctx, cudaMalloc(&pl->admm_c, nreg * pl->esz));
        pl->bytes += nreg * pl->esz;
    }
    NSOL_CHECK(lsq_ensure_stage(pl, nreg * sizeof(double)));
    NSOL_CUDA(ctx, cudaMemcpyAsync(pl->stage, b_reg_host, nreg * sizeof(double), cudaMemcpyHostToDevice, st));
    NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)nreg, NSOL_F64, pl->stage, pl->gv.dtype, pl->admm_c, in_scale, 1, st));
    NSOL_CUDA(ctx, cudaStreamSynchronize(st));      // b_reg_host may be a temporary
    return NSOL_OK;
}

extern "C" int nsol_admm_shrink(nsol_ctx *ctx, const nsol_grid *grid, const void *x_dev, const void *w_in_dev, double ell, void *v_dev,
                                void *w_dev, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    if (!x_dev || !v_dev || !w_dev) return nsol_fail(ctx, NSOL_EINVAL, "nsol_admm_shrink: NULL array");
    NSOL_CHECK(nsol_bind_device(ctx));
    nsol_lsmr_plan tmp;
    tmp.ctx = ctx;
    tmp.desc.b_op = NSOL_B_GRAD;
    NSOL_CHECK(nsol_grid_view(ctx, grid, &tmp.gv));
    if (tmp.gv.batch != 1) return nsol_fail(ctx, NSOL_EINVAL, "nsol_admm_shrink: grid.batch must be 1");
    long long want = (tmp.gv.n + LSMR_THREADS - 1) / LSMR_THREADS;
    long long cap = (long long)ctx->sm_count * 8;
    const int nb = (int)(want < cap ? (want > 0 ? want : 1) : cap);
    if (tmp.gv.dtype == NSOL_F32)
        admm_shrink_kernel<float><<<nb, LSMR_THREADS, 0, (cudaStream_t)s>>>(make_geom<float>(&tmp), (const float *)x_dev, (const float *)w_in_dev,
                                                                           (float)ell, (float *)v_dev, (float *)w_dev, (float *)nullptr);
    else
        admm_shrink_kernel<double><<<nb, LSMR_THREADS, 0, (cudaStream_t)s>>>(make_geom<double>(&tmp), (const double *)x_dev, (const double *)w_in_dev,
                                                                            ell, (double *)v_dev, (double *)w_dev, (double *)nullptr);
    NSOL_LAUNCH_CHECK(ctx);
    return NSOL_OK;
}

// ---------------------------------------------------------------------------
// z-slab decomposition of the LSMR / ADMM path (one rank per GPU)
// ---------------------------------------------------------------------------
// The plan's volume is the local slab [z_lo, z_hi) of a taller volume.  Communication is the caller's
// (nsol_b200/distributed.py: NCCL send/recv between neighbours + all-reduce of one double); this side
// exposes the planes to send, the halo buffers to receive into, and the solve cut into phases at every
// point where a neighbour exchange or a global sum is needed.  Blur halos are periodic (ring: rank 0's
// lower neighbour is the last rank); gradient stencils keep the zero boundary at the global ends.
__global__ void lsmr_reduce_ss_kernel(const LsmrScalars *__restrict__ S, const double *__restrict__ part, int count, double *__restrict__ ss,
                                      int always) {
    if (!always && S->done) return;
    const double v = reduce_partials(part, count);
    if (threadIdx.x == 0) ss[0] = v;
}

extern "C" int nsol_lsmr_plan_slab(nsol_lsmr_plan *pl, int has_below, int has_above) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    const GridView &gv = pl->gv;
    if (gv.dim < 2) return nsol_fail(ctx, NSOL_EINVAL, "lsmr slab: needs dim >= 2");
    NSOL_CHECK(nsol_bind_device(ctx));
    int ghost = 1;
    if (pl->desc.a_op == NSOL_A_BLUR && pl->desc.radius[0] > ghost) ghost = pl->desc.radius[0];   // numpy axis 0 = slab axis
    if (gv.nz < ghost) return nsol_fail(ctx, NSOL_EINVAL, "lsmr slab: %d local planes but the blur needs %d halo planes", gv.nz, ghost);
    if (!pl->slab) {
        const size_t plane = (size_t)gv.nx * gv.ny * pl->esz;
        void **bufs[] = {&pl->halo_v_lo, &pl->halo_v_hi, &pl->halo_u_lo, &pl->halo_u_hi, &pl->halo_uz_lo, &pl->halo_x_hi};
        const size_t sizes[] = {plane * ghost, plane * ghost, plane * ghost, plane * ghost, plane, plane};
        for (int i = 0; i < 6; ++i) {
            NSOL_CUDA(ctx, nsol_plan_alloc(ctx, bufs[i], sizes[i]));
            NSOL_CUDA(ctx, cudaMemset(*bufs[i], 0, sizes[i]));
            pl->bytes += sizes[i];
        }
        NSOL_CUDA(ctx, cudaMalloc((void **)&pl->ssbuf, sizeof(double)));
    }
    pl->slab = true;
    pl->ghost = ghost;
    pl->grad_lo = has_below ? 1 : 0;
    pl->grad_hi = has_above ? 1 : 0;
    return NSOL_OK;
}

extern "C" int nsol_lsmr_slab_buffers(nsol_lsmr_plan *pl, int which, const void **send_first, const void **send_last, void **recv_lo,
                                      void **recv_hi, int *planes) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!pl->slab) return nsol_fail(ctx, NSOL_ESTATE, "lsmr slab: call nsol_lsmr_plan_slab first");
    const GridView &gv = pl->gv;
    const size_t plane = (size_t)gv.nx * gv.ny * pl->esz;
    const void *sf = nullptr, *sl = nullptr;
    void *rl = nullptr, *rh = nullptr;
    int np = 0;
    auto first_last = [&](const char *base, int count) {
        sf = base;
        sl = base + (size_t)(gv.nz - count) * plane;
        np = count;
    };
    switch (which) {
    case NSOL_SLAB_V:       // vhat: ghost planes both ways (periodic blur); hi plane 0 also feeds grad
        first_last((const char *)pl->v, pl->ghost);
        rl = pl->halo_v_lo;
        rh = pl->halo_v_hi;
        break;
    case NSOL_SLAB_U0:      // u block 0: ghost planes both ways (A^T)
        first_last((const char *)pl->u, pl->ghost);
        rl = pl->halo_u_lo;
        rh = pl->halo_u_hi;
        break;
    case NSOL_SLAB_UZ:      // u block of D_z (the last B block): my last plane -> upper neighbour's lo buffer
        first_last((const char *)pl->u + (size_t)gv.dim * gv.n * pl->esz, 1);
        sf = nullptr;
        rl = pl->halo_uz_lo;
        break;
    case NSOL_SLAB_X:       // x: my first plane -> lower neighbour's hi buffer
        first_last((const char *)pl->xbuf, 1);
        sl = nullptr;
        rh = pl->halo_x_hi;
        break;
    default:
        return nsol_fail(ctx, NSOL_EINVAL, "lsmr slab: unknown buffer kind %d", which);
    }
    if (send_first) *send_first = sf;
    if (send_last) *send_last = sl;
    if (recv_lo) *recv_lo = rl;
    if (recv_hi) *recv_hi = rh;
    if (planes) *planes = np;
    return NSOL_OK;
}

extern "C" int nsol_lsmr_slab_arrays(nsol_lsmr_plan *pl, void **b_dev, void **x_dev, double **ss_dev) {
    if (!pl) return NSOL_EINVAL;
    if (!pl->slab) return nsol_fail(pl->ctx, NSOL_ESTATE, "lsmr slab: call nsol_lsmr_plan_slab first");
    if (b_dev) *b_dev = pl->bbuf;
    if (x_dev) *x_dev = pl->xbuf;
    if (ss_dev) *ss_dev = pl->ssbuf;
    return NSOL_OK;
}

template <typename T>
static int lsmr_slab_phase_t(nsol_lsmr_plan *pl, int phase, double p0, double p1, int i0, cudaStream_t s) {
    nsol_ctx *ctx = pl->ctx;
    const LsqGeom<T> g = make_geom<T>(pl);
    double *part = pl->part;
    if (g.ny > 65535 || g.nz > 65535) return nsol_fail(ctx, NSOL_EINVAL, "lsmr slab: more than 65535 rows along y or z are not supported");
    int rparts = 0;
    switch (phase) {
    case NSOL_PH_RHS:            // u = [b; sqrt_alpha b_reg]; p0 = sqrt_alpha, i0 = 0: b_reg = 0  -> ss
        NSOL_CHECK(lsq_launch_rhs<T>(pl, g, pl->rows_b, pl->bbuf, i0 ? pl->breg : nullptr, p0, s, &rparts));
        lsmr_reduce_ss_kernel<<<1, 1024, 0, s>>>(pl->S, part, rparts, pl->ssbuf, 1);
        break;
    case NSOL_PH_SCAL_INIT_BETA:  // p0 = sqrt_alpha, i0 = maxiter
        lsmr_scalar_init_beta<<<1, 32, 0, s>>>(pl->S, pl->ssbuf, 1, p0, i0);
        break;
    case NSOL_PH_ADJ_FIRST:
    case NSOL_PH_ADJ:            // needs the U0 and UZ halos -> ss
        NSOL_CHECK(lsq_launch_adj<T>(pl, g, pl->halo_u_lo, pl->halo_u_hi, phase == NSOL_PH_ADJ_FIRST ? 1 : 0, pl->halo_uz_lo, s, &rparts));
        lsmr_reduce_ss_kernel<<<1, 1024, 0, s>>>(pl->S, part, rparts, pl->ssbuf, phase == NSOL_PH_ADJ_FIRST ? 1 : 0);
        break;
    case NSOL_PH_SCAL_INIT_ALPHA:
        lsmr_scalar_init_alpha<<<1, 32, 0, s>>>(pl->S, pl->ssbuf, 1);
        NSOL_LAUNCH_CHECK(ctx);
        NSOL_CHECK(lsq_launch_init_vectors<T>(pl, g, s));
        return NSOL_OK;
    case NSOL_PH_FWD:            // needs the V halos -> ss
        NSOL_CHECK(lsq_launch_fwd<T>(pl, g, pl->halo_v_lo, pl->halo_v_hi, pl->halo_v_hi, s, &rparts));
        lsmr_reduce_ss_kernel<<<1, 1024, 0, s>>>(pl->S, part, rparts, pl->ssbuf, 0);
        break;
    case NSOL_PH_SCAL_BETA:
        lsmr_scalar_beta<<<1, 32, 0, s>>>(pl->S, pl->ssbuf, 1);
        break;
    case NSOL_PH_SCAL_ALPHA:
        lsmr_scalar_alpha<<<1, 32, 0, s>>>(pl->S, pl->ssbuf, 1);
        break;
    case NSOL_PH_UPDATE:         // -> ss
        NSOL_CHECK(lsq_launch_update<T>(pl, g, s, &rparts));
        lsmr_reduce_ss_kernel<<<1, 1024, 0, s>>>(pl->S, part, rparts, pl->ssbuf, 0);
        break;
    case NSOL_PH_SCAL_TESTS:
        lsmr_scalar_tests<<<1, 32, 0, s>>>(pl->S, pl->ssbuf, 1);
        break;
    case NSOL_PH_CLIP:           // xbuf = clip(x, p0, p1)
        NSOL_CHECK(lsq_launch_clip<T>(pl, g, pl->xbuf, p0, p1, s));
        return NSOL_OK;
    case NSOL_PH_ADMM_INIT:      // needs the X halo: v = grad(xbuf), w = 0, b_reg = v
        NSOL_CHECK(lsq_launch_shrink<T>(pl, g, pl->xbuf, nullptr, 0.0, pl->admm_v, pl->admm_w, pl->breg, pl->halo_x_hi, 1, s));
        return NSOL_OK;
    case NSOL_PH_ADMM_SHRINK:    // needs the X halo: p0 = ell = alpha / rho
        NSOL_CHECK(lsq_launch_shrink<T>(pl, g, pl->xbuf, pl->admm_w, p0, pl->admm_v, pl->admm_w, pl->breg, pl->halo_x_hi, 0, s));
        return NSOL_OK;
    default:
        return nsol_fail(ctx, NSOL_EINVAL, "lsmr slab: unknown phase %d", phase);
    }
    NSOL_LAUNCH_CHECK(ctx);
    return NSOL_OK;
}

extern "C" int nsol_lsmr_slab_phase(nsol_lsmr_plan *pl, int phase, double p0, double p1, int i0, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    if (!pl->slab) return nsol_fail(pl->ctx, NSOL_ESTATE, "lsmr slab: call nsol_lsmr_plan_slab first");
    NSOL_CHECK(nsol_bind_device(pl->ctx));
    if (pl->gv.dtype == NSOL_F32) return lsmr_slab_phase_t<float>(pl, phase, p0, p1, i0, (cudaStream_t)s);
    return lsmr_slab_phase_t<double>(pl, phase, p0, p1, i0, (cudaStream_t)s);
}

// current LSMR state of a plan (synchronises): iteration count and stop flag
extern "C" int nsol_lsmr_plan_status(nsol_lsmr_plan *pl, int *itn_out, int *istop_out, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    NSOL_CHECK(nsol_bind_device(ctx));
    LsmrScalars hS;
    NSOL_CUDA(ctx, cudaMemcpyAsync(&hS, pl->S, sizeof(hS), cudaMemcpyDeviceToHost, (cudaStream_t)s));
    NSOL_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)s));
    if (itn_out) *itn_out = hS.itn;
    if (istop_out) *istop_out = hS.istop;
    return NSOL_OK;
}

// ---------------------------------------------------------------------------
// primal-dual deconvolution: prox_f = prox_linear_least_squares (one LSMR solve per PD iteration)
// ---------------------------------------------------------------------------
// p <- prox_g*(p + sigma grad(xbar))   (primal_dual_solver.py:242-243; proximal_operators.py:139-140, 157-159)
// Step sizes: row *it of the device table sched[iterations][8] = (sigma, tau, tau*lambda, theta, den_g, den_f, 0, 0)
// (pd_schedule_rows, float64 host arithmetic in the reference's order) -- read on the device so that ONE captured CUDA graph
// of a primal-dual iteration can be replayed for every iteration.
template <typename T>
__global__ void pdd_dual_kernel(LsqGeom<T> g, const T *__restrict__ xbar, T *__restrict__ p, const double *__restrict__ sched,
                                const int *__restrict__ it, int reg) {
    const double *row = sched + (long long)(*it) * 8;
    const T sigma = (T)row[0], den = (T)row[4];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += (long long)gridDim.x * blockDim.x) {
        int idx[3];
        lsq_decode(g, i, idx);
        const T xc = xbar[i];
        for (int k = 0; k < g.dim; ++k) {
            const T hi = (idx[g.axis[k]] + 1 < g.extent[k]) ? xbar[i + g.stride[k]] : T(0);
            const T gk = g.w[k] * hi + (-g.w[k]) * xc;
            T q = p[(long long)k * g.n + i] + sigma * gk;
            if (reg != NSOL_REG_TV) q = q / den;
            if (reg != NSOL_REG_TK1) q = q / max_t(T(1), abs_t(q));
            p[(long long)k * g.n + i] = q;
        }
    }
}

// b_reg <- (x - tau grad_adj(p)) / prox_scale   (primal_dual_solver.py:246; tikhonov b_reg / x_scale)
template <typename T>
__global__ void pdd_arg_kernel(LsqGeom<T> g, const T *__restrict__ x, const T *__restrict__ p, const double *__restrict__ sched,
                               const int *__restrict__ it, T prox_scale, T *__restrict__ breg) {
    const T tau = (T)sched[(long long)(*it) * 8 + 1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += (long long)gridDim.x * blockDim.x) {
        int idx[3];
        lsq_decode(g, i, idx);
        T div = T(0);
        for (int k = 0; k < g.dim; ++k) {
            const T *pk = p + (long long)k * g.n;
            const T lo = (idx[g.axis[k]] > 0) ? pk[i - g.stride[k]] : T(0);
            const T dk = g.w[k] * lo + (-g.w[k]) * pk[i];
            div = (k == 0) ? dk : div + dk;
        }
        breg[i] = (x[i] - tau * div) / prox_scale;
    }
}

// x+ = y * prox_scale ; xbar = x+ + theta (x+ - x) ; x = x+   (solver.py:117-118; primal_dual_solver.py:253)
template <typename T>
__global__ void pdd_relax_kernel(long long n, const T *__restrict__ y, T prox_scale, const double *__restrict__ sched,
                                 const int *__restrict__ it, T *__restrict__ x, T *__restrict__ xbar) {
    const T theta = (T)sched[(long long)(*it) * 8 + 3];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T xn = y[i] * prox_scale;
        const T xo = x[i];
        xbar[i] = xn + theta * (xn - xo);
        x[i] = xn;
    }
}

// device-side bookkeeping of the graph-replayed iteration: sa = sqrt(1 / (tau * lambda)) of the current row (the weight of
// the identity rows of the prox's Tikhonov system, proximal_operators.py:58-75) before the solve; ++it after the relaxation
__global__ void pdd_prepare_kernel(const double *__restrict__ sched, const int *__restrict__ it, double *__restrict__ sa) {
    sa[0] = sqrt(1.0 / sched[(long long)(*it) * 8 + 2]);
}
__global__ void pdd_advance_kernel(int *it) { it[0] += 1; }

template <typename T>
static int pd_deconv_t(nsol_lsmr_plan *pl, const nsol_pd_desc *pd, int iterations, int iter_max, double prox_scale, double *x_host,
                       double *iterates_host, cudaStream_t s) {
    nsol_ctx *ctx = pl->ctx;
    LsqGeom<T> g = make_geom<T>(pl);
    const int nb = pl->nblocks, th = LSMR_THREADS;
    const size_t n = (size_t)pl->gv.n;
    constexpr int VEC = FastvCfg<T>::VEC;
    const bool vec = fastv_ok(pl) && g.ny <= 65535 && g.nz <= 65535;      // row-mapped 128-bit kernels (csrc/lsmr_fastv.cuh)
    const FastvGeom<T> fg = make_fastv_geom<T>(g);
    const dim3 vgrid((g.nx / VEC + FAST_TH - 1) / FAST_TH, g.ny, g.nz);
    // state: x = pl->xbuf; admm_v holds [xbar | y | spare], admm_w holds p (the LSMR solve owns u, v, h, hbar, x)
    T *x = (T *)pl->xbuf, *xbar = (T *)pl->admm_v, *p = (T *)pl->admm_w, *breg = (T *)pl->breg;
    T *y = (pl->gv.dim >= 2) ? xbar + n : nullptr;
    void *ybuf = y;
    bool own_y = false;
    if (!ybuf) {
        NSOL_CUDA(ctx, cudaMalloc(&ybuf, n * sizeof(T)));
        own_y = true;
    }
    // step-size table on the device: [iterations][8] doubles, then sa (1 double) and the iteration counter
    const int rows_n = iterations > 0 ? iterations : 1;
    std::vector<double> rows((size_t)rows_n * 8);
    pd_schedule_rows(*pd, pd->alpha[0], iterations, rows.data());
    for (int i = 0; i < iterations; ++i)
        if (!(1.0 / rows[(size_t)i * 8 + 2] > 1e-10)) {
            if (own_y) cudaFree(ybuf);
            return nsol_fail(ctx, NSOL_EINVAL, "pd deconvolution: tau * lambda = %g is too large for the Tikhonov weight", rows[(size_t)i * 8 + 2]);
        }
    char *tab = nullptr;
    const size_t tab_bytes = rows.size() * sizeof(double) + 2 * sizeof(double);
    cudaError_t e0 = cudaMalloc((void **)&tab, tab_bytes);
    if (e0 != cudaSuccess) {
        if (own_y) cudaFree(ybuf);
        return nsol_fail(ctx, NSOL_ENOMEM, "pd deconvolution: cudaMalloc(%zu) -> %s", tab_bytes, cudaGetErrorString(e0));
    }
    double *sched = (double *)tab, *sa_dev = sched + rows.size();
    int *it_dev = (int *)(sa_dev + 1);
    auto cleanup = [&]() {
        cudaStreamSynchronize(s);
        cudaFree(tab);
        if (own_y) cudaFree(ybuf);
    };
    int rc = NSOL_OK;
    {
        cudaError_t e = cudaMemcpyAsync(sched, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(sa_dev, 0, 2 * sizeof(double), s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(xbar, x, n * sizeof(T), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, n * pl->gv.dim * sizeof(T), s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);       // `rows` is a host temporary
        if (e != cudaSuccess) rc = nsol_fail(ctx, NSOL_ECUDA, "pd deconvolution: %s", cudaGetErrorString(e));
    }
    if (rc == NSOL_OK && iterates_host) rc = lsq_download(pl, x, pd->x_scale, iterates_host, s);
    const bool coopv = coopv_ok(pl);
    const int coop = (rc == NSOL_OK && !coopv) ? lsmr_use_coop(pl) : 0;
    if (coop < 0) rc = coop;
    const int reg = pd->reg;

    // one primal-dual iteration (primal_dual_solver.py:242-253); every kernel takes its step sizes from row *it_dev
    auto iteration = [&](cudaStream_t st, double alpha_host) -> int {
        if (vec) {
            fastv_pdd_dual_kernel<T, VEC><<<vgrid, FAST_TH, 0, st>>>(fg, xbar, p, sched, it_dev, reg);
            NSOL_LAUNCH_CHECK(ctx);
            fastv_pdd_arg_kernel<T, VEC><<<vgrid, FAST_TH, 0, st>>>(fg, x, p, sched, it_dev, (T)prox_scale, breg);
            NSOL_LAUNCH_CHECK(ctx);
        } else {
            pdd_dual_kernel<T><<<nb, th, 0, st>>>(g, xbar, p, sched, it_dev, reg);
            NSOL_LAUNCH_CHECK(ctx);
            pdd_arg_kernel<T><<<nb, th, 0, st>>>(g, x, p, sched, it_dev, (T)prox_scale, breg);
            NSOL_LAUNCH_CHECK(ctx);
        }
        // tikhonov: alpha = 1 / (tau * lambda), b_reg = y / prox_scale, B = I  (proximal_operators.py:58-75)
        if (coopv) {
            NSOL_CHECK(lsmr_solve_any(pl, alpha_host, pl->bbuf, breg, iter_max, 0.0, INFINITY, ybuf, st));
        } else if (coop) {
            NSOL_CHECK(lsmr_solve_coop<T>(pl, alpha_host, pl->bbuf, breg, iter_max, 0.0, INFINITY, ybuf, 0, 0.0, st));
        } else {
            pdd_prepare_kernel<<<1, 1, 0, st>>>(sched, it_dev, sa_dev);
            NSOL_LAUNCH_CHECK(ctx);
            NSOL_CHECK(lsmr_solve_t<T>(pl, 1.0, pl->bbuf, breg, iter_max, 0.0, INFINITY, ybuf, st, sa_dev));
        }
        if (vec) fastv_pdd_relax_kernel<T, VEC><<<lsq_flat_blocks(pl, (long long)n / VEC), th, 0, st>>>((long long)n / VEC, (const T *)ybuf, (T)prox_scale, sched, it_dev, x, xbar);
        else pdd_relax_kernel<T><<<nb, th, 0, st>>>((long long)n, (const T *)ybuf, (T)prox_scale, sched, it_dev, x, xbar);
        NSOL_LAUNCH_CHECK(ctx);
        pdd_advance_kernel<<<1, 1, 0, st>>>(it_dev);
        NSOL_LAUNCH_CHECK(ctx);
        return NSOL_OK;
    };

    if (rc == NSOL_OK && !coop && !coopv && !iterates_host && iterations >= 2) {
        // The launch sequence of an iteration is fixed and all its parameters live on the device: capture it once,
        // replay it `iterations` times (~90 kernels per iteration stop paying individual launch latency).
        const int64_t l0 = ctx->launches;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        if (ce == cudaSuccess) {
            rc = iteration(s, 0.0);
            ce = cudaStreamEndCapture(s, &graph);
        }
        if (rc == NSOL_OK && ce != cudaSuccess) rc = nsol_fail(ctx, NSOL_ECUDA, "pd deconvolution: graph capture failed: %s", cudaGetErrorString(ce));
        const int64_t per_iteration = ctx->launches - l0;
        if (rc == NSOL_OK) {
            ce = cudaGraphInstantiate(&exec, graph, 0);
            if (ce != cudaSuccess) rc = nsol_fail(ctx, NSOL_ECUDA, "pd deconvolution: graph instantiation failed: %s", cudaGetErrorString(ce));
        }
        if (graph) cudaGraphDestroy(graph);
        for (int it = 0; rc == NSOL_OK && it < iterations; ++it) {
            ce = cudaGraphLaunch(exec, s);
            if (ce != cudaSuccess) rc = nsol_fail(ctx, NSOL_ECUDA, "pd deconvolution: graph launch failed: %s", cudaGetErrorString(ce));
        }
        ctx->launches += per_iteration * (iterations - 1);
        if (exec) {
            cudaError_t se = cudaStreamSynchronize(s);       // the executable graph must outlive its launches
            cudaGraphExecDestroy(exec);
            if (rc == NSOL_OK && se != cudaSuccess) rc = nsol_fail(ctx, NSOL_ECUDA, "pd deconvolution: %s", cudaGetErrorString(se));
        }
    } else {
        for (int it = 0; it < iterations && rc == NSOL_OK; ++it) {
            rc = iteration(s, 1.0 / rows[(size_t)it * 8 + 2]);
            if (rc == NSOL_OK && iterates_host) rc = lsq_download(pl, x, pd->x_scale, iterates_host + (size_t)(it + 1) * n, s);
        }
    }
    if (rc == NSOL_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = nsol_fail(ctx, NSOL_ECUDA, "pd deconvolution: %s", cudaGetErrorString(e));
    }
    if (rc == NSOL_OK) rc = lsq_download(pl, x, pd->x_scale, x_host, s);
    cleanup();
    return rc;
}

extern "C" int nsol_pd_deconv_run_host(nsol_lsmr_plan *pl, const nsol_pd_desc *pd, int iterations, int iter_max, double prox_scale,
                                       const double *b_host, const double *x0_host, double *x_host, double *iterates_host, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!pd || !b_host || !x0_host || !x_host || !pd->alpha) return nsol_fail(ctx, NSOL_EINVAL, "pd deconvolution: NULL argument");
    if (pl->desc.b_op != NSOL_B_IDENTITY) return nsol_fail(ctx, NSOL_EINVAL, "pd deconvolution: the plan's B must be the identity");
    if (iterations < 0 || iter_max < 0) return nsol_fail(ctx, NSOL_EINVAL, "pd deconvolution: iterations and iter_max must be >= 0");
    if (prox_scale == 0.0 || pd->x_scale == 0.0 || !(pd->alpha[0] > 0.0) || !(pd->L2 > 0.0))
        return nsol_fail(ctx, NSOL_EINVAL, "pd deconvolution: scales, alpha and L2 must be non-zero / positive");
    if (pd->reg != NSOL_REG_TV && pd->reg != NSOL_REG_HUBER && pd->reg != NSOL_REG_TK1) return nsol_fail(ctx, NSOL_EINVAL, "pd deconvolution: unknown regulariser");
    if (pd->alg < NSOL_ALG2 || pd->alg > NSOL_ALG3) return nsol_fail(ctx, NSOL_EINVAL, "pd deconvolution: unknown alg_type");
    GridView gv;
    NSOL_CHECK(nsol_grid_view(ctx, &pd->grid, &gv));
    if (gv.dim != pl->gv.dim || gv.nx != pl->gv.nx || gv.ny != pl->gv.ny || gv.nz != pl->gv.nz || gv.dtype != pl->gv.dtype || gv.batch != 1)
        return nsol_fail(ctx, NSOL_EINVAL, "pd deconvolution: pd.grid must equal the plan's grid (batch 1)");
    NSOL_CHECK(nsol_bind_device(ctx));
    cudaStream_t st = nullptr;
    NSOL_CHECK(admm_stream(pl, s, &st));
    const size_t n = (size_t)pl->gv.n;
    NSOL_CHECK(lsq_ensure_stage(pl, 2 * n * sizeof(double)));
    double *sb = (double *)pl->stage;
    NSOL_CUDA(ctx, cudaMemcpyAsync(sb, b_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    NSOL_CUDA(ctx, cudaMemcpyAsync(sb + n, x0_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    // the prox hands b / prox_scale to a solver that divides by prox_scale again (proximal_operators.py:62, linear_solver.py:83)
    NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)n, NSOL_F64, sb, NSOL_F64, sb, prox_scale, 1, st));
    NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)n, NSOL_F64, sb, pl->gv.dtype, pl->bbuf, prox_scale, 1, st));
    NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)n, NSOL_F64, sb + n, pl->gv.dtype, pl->xbuf, pd->x0_scale, 1, st));
    if (pl->gv.dtype == NSOL_F32) return pd_deconv_t<float>(pl, pd, iterations, iter_max, prox_scale, x_host, iterates_host, st);
    return pd_deconv_t<double>(pl, pd, iterations, iter_max, prox_scale, x_host, iterates_host, st);
}
