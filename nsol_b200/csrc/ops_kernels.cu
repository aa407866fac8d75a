// ops_kernels.cu -- standalone linear operators on device arrays:
// zero-boundary forward-difference gradient and its adjoint (reference
// nsol/linear_operators.py:98-169), periodic separable / dense convolution
// (:60-86).  These are the building blocks the LSMR/ADMM path fuses further
// (lsmr_kernels.cu); on their own they back the LinearOperators* closures.
#include <vector>

#include "common.cuh"

// ---------------------------------------------------------------------------
// gradient / single differences
// ---------------------------------------------------------------------------
// D_k x [i] = fl(fl(w x[i+e_k]) + fl((-w) x[i])), x beyond the upper end := 0
// D_k^T y[i] = fl(fl(w y[i-e_k]) + fl((-w) y[i])), y below the lower end := 0
template <typename T>
struct GradArgs {
    const T *in;     // x (grad) or p (adjoint)
    T *out;          // p (grad) or x (adjoint)
    long long n;     // voxels per problem
    long long total; // n * batch
    int nx, ny, nz;
    int dim;
    // stride (in elements) and extent of reference component k
    long long stride[3];
    int extent[3];
    T w[3];
};

template <typename T>
__device__ __forceinline__ void decode_axes(long long r, int nx, int ny, int &ix, int &iy, int &iz) {
    ix = (int)(r % nx);
    long long t = r / nx;
    iy = (int)(t % ny);
    iz = (int)(t / ny);
}

template <typename T>
__global__ void grad_kernel(const GradArgs<T> a) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < a.total; i += step) {
        const long long bi = i / a.n, r = i - bi * a.n;
        int idx[3];
        decode_axes<T>(r, a.nx, a.ny, idx[0], idx[1], idx[2]);
        // component k -> kernel axis: k=0 -> x; 3-D: k=1 -> y, k=2 -> z; 2-D: k=1 -> z
        const T c = a.in[i];
        for (int k = 0; k < a.dim; ++k) {
            const int ax = (k == 0) ? 0 : ((a.dim == 3 && k == 1) ? 1 : 2);
            const T hi = (idx[ax] + 1 < a.extent[k]) ? a.in[i + a.stride[k]] : T(0);
            a.out[(bi * a.dim + k) * a.n + r] = a.w[k] * hi + (-a.w[k]) * c;
        }
    }
}

template <typename T>
__global__ void grad_adj_kernel(const GradArgs<T> a) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < a.total; i += step) {
        const long long bi = i / a.n, r = i - bi * a.n;
        int idx[3];
        decode_axes<T>(r, a.nx, a.ny, idx[0], idx[1], idx[2]);
        T acc = T(0);
        for (int k = 0; k < a.dim; ++k) {
            const int ax = (k == 0) ? 0 : ((a.dim == 3 && k == 1) ? 1 : 2);
            const T *pk = a.in + (bi * a.dim + k) * a.n;
            const T lo = (idx[ax] > 0) ? pk[r - a.stride[k]] : T(0);
            const T d = a.w[k] * lo + (-a.w[k]) * pk[r];
            acc = (k == 0) ? d : acc + d;   // (Dx^T + Dy^T) + Dz^T  (linear_operators.py:166-168)
        }
        a.out[i] = acc;
    }
}

// single component: out = D_k in  or  D_k^T in  (both N -> N)
template <typename T, bool ADJ>
__global__ void diff_kernel(const GradArgs<T> a, int k) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    const int ax = (k == 0) ? 0 : ((a.dim == 3 && k == 1) ? 1 : 2);
    for (; i < a.total; i += step) {
        const long long bi = i / a.n, r = i - bi * a.n;
        int idx[3];
        decode_axes<T>(r, a.nx, a.ny, idx[0], idx[1], idx[2]);
        T nb;
        if (ADJ) nb = (idx[ax] > 0) ? a.in[i - a.stride[k]] : T(0);
        else nb = (idx[ax] + 1 < a.extent[k]) ? a.in[i + a.stride[k]] : T(0);
        a.out[i] = a.w[k] * nb + (-a.w[k]) * a.in[i];
    }
}

template <typename T>
static void fill_grad_args(const GridView &gv, const void *in, void *out, GradArgs<T> &a) {
    a.in = (const T *)in;
    a.out = (T *)out;
    a.n = gv.n;
    a.total = gv.n * gv.batch;
    a.nx = gv.nx;
    a.ny = gv.ny;
    a.nz = gv.nz;
    a.dim = gv.dim;
    for (int k = 0; k < 3; ++k) {
        a.stride[k] = 0;
        a.extent[k] = 1;
        a.w[k] = T(0);
    }
    for (int k = 0; k < gv.dim; ++k) {
        const int ax = (k == 0) ? 0 : ((gv.dim == 3 && k == 1) ? 1 : 2);
        a.stride[k] = ax == 0 ? 1 : (ax == 1 ? gv.nx : (long long)gv.nx * gv.ny);
        a.extent[k] = ax == 0 ? gv.nx : (ax == 1 ? gv.ny : gv.nz);
        a.w[k] = (T)gv.w[k];
    }
}

static int grid_1d(const nsol_ctx *ctx, long long total, int threads) {
    long long want = (total + threads - 1) / threads;
    long long cap = (long long)ctx->sm_count * 32;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

template <typename T>
static int launch_grad_t(nsol_ctx *ctx, const GridView &gv, int mode, int k, const void *in, void *out, cudaStream_t s) {
    GradArgs<T> a;
    fill_grad_args<T>(gv, in, out, a);
    const int threads = 256;
    const int blocks = grid_1d(ctx, a.total, threads);
    if (mode == 0) grad_kernel<T><<<blocks, threads, 0, s>>>(a);
    else if (mode == 1) grad_adj_kernel<T><<<blocks, threads, 0, s>>>(a);
    else if (mode == 2) diff_kernel<T, false><<<blocks, threads, 0, s>>>(a, k);
    else diff_kernel<T, true><<<blocks, threads, 0, s>>>(a, k);
    NSOL_LAUNCH_CHECK(ctx);
    return NSOL_OK;
}

static int launch_grad(nsol_ctx *ctx, const nsol_grid *g, int mode, int k, const void *in, void *out, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    if (!in || !out) return nsol_fail(ctx, NSOL_EINVAL, "operator: NULL array");
    NSOL_CHECK(nsol_bind_device(ctx));
    GridView gv;
    NSOL_CHECK(nsol_grid_view(ctx, g, &gv));
    if (mode >= 2 && (k < 0 || k >= gv.dim)) return nsol_fail(ctx, NSOL_EINVAL, "nsol_diff: component %d out of range for dim %d", k, gv.dim);
    if (gv.dtype == NSOL_F32) return launch_grad_t<float>(ctx, gv, mode, k, in, out, (cudaStream_t)s);
    return launch_grad_t<double>(ctx, gv, mode, k, in, out, (cudaStream_t)s);
}

extern "C" int nsol_grad(nsol_ctx *ctx, const nsol_grid *g, const void *x_dev, void *p_dev, nsol_stream s) {
    return launch_grad(ctx, g, 0, 0, x_dev, p_dev, s);
}
extern "C" int nsol_grad_adj(nsol_ctx *ctx, const nsol_grid *g, const void *p_dev, void *x_dev, nsol_stream s) {
    return launch_grad(ctx, g, 1, 0, p_dev, x_dev, s);
}
extern "C" int nsol_diff(nsol_ctx *ctx, const nsol_grid *g, int component, int adjoint, const void *in_dev, void *out_dev, nsol_stream s) {
    return launch_grad(ctx, g, adjoint ? 3 : 2, component, in_dev, out_dev, s);
}

// ---------------------------------------------------------------------------
// periodic separable convolution: one pass per numpy axis (axis 0 first)
// ---------------------------------------------------------------------------
#define NSOL_MAX_TAPS 129

template <typename T>
struct ConvAxisArgs {
    const T *in;
    T *out;
    long long total;   // elements over all problems
    long long stride;  // element stride of the axis
    int extent;        // axis length
    int radius;
    T taps[NSOL_MAX_TAPS];
};

// out[i] = sum_k taps[k] * in[(i - (k - r)) mod n]   (scipy.ndimage.convolve, mode="wrap")
template <typename T>
__global__ void conv_axis_wrap_kernel(const __grid_constant__ ConvAxisArgs<T> a) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    const bool small = a.total <= 0x7fffffffLL;          // 32-bit index arithmetic
    const bool narrow = a.radius < a.extent;              // wrap with one conditional add / subtract
    for (; i < a.total; i += step) {
        const int pos = small ? (int)(((unsigned)i / (unsigned)a.stride) % (unsigned)a.extent) : (int)((i / a.stride) % a.extent);
        const T *line = a.in + (i - (long long)pos * a.stride);
        T acc = T(0);
        if (narrow) {
#pragma unroll 7
            for (int k = 0; k <= 2 * a.radius; ++k) {
                int q = pos - (k - a.radius);
                q += q < 0 ? a.extent : 0;
                q -= q >= a.extent ? a.extent : 0;
                acc += a.taps[k] * line[(long long)q * a.stride];
            }
        } else {
            for (int k = 0; k <= 2 * a.radius; ++k) {
                int q = (pos - (k - a.radius)) % a.extent;
                if (q < 0) q += a.extent;
                acc += a.taps[k] * line[(long long)q * a.stride];
            }
        }
        a.out[i] = acc;
    }
}

template <typename T>
static int blur_sep_t(nsol_ctx *ctx, const GridView &gv, const nsol_grid *g, const double *const taps[3], const int32_t radius[3],
                      const void *x, void *y, void *tmp, cudaStream_t s) {
    // numpy axis a of a dim-D array -> element stride
    const int dim = gv.dim;
    long long strides[3];
    long long acc = 1;
    for (int a = dim - 1; a >= 0; --a) {
        strides[a] = acc;
        acc *= g->shape[a];
    }
    // ping-pong so that the last pass lands in y
    const void *src = x;
    for (int a = 0; a < dim; ++a) {
        const int remaining = dim - 1 - a;
        void *dst = (remaining % 2 == 0) ? y : tmp;
        if (!dst) return nsol_fail(ctx, NSOL_EINVAL, "nsol_blur_sep: tmp_dev is required for dim >= 2");
        ConvAxisArgs<T> c;
        c.in = (const T *)src;
        c.out = (T *)dst;
        c.total = gv.n * gv.batch;
        c.stride = strides[a];
        c.extent = (int)g->shape[a];
        c.radius = radius[a];
        for (int k = 0; k <= 2 * radius[a]; ++k) c.taps[k] = (T)taps[a][k];
        const int threads = 256;
        conv_axis_wrap_kernel<T><<<grid_1d(ctx, c.total, threads), threads, 0, s>>>(c);
        NSOL_LAUNCH_CHECK(ctx);
        src = dst;
    }
    return NSOL_OK;
}

extern "C" int nsol_blur_sep(nsol_ctx *ctx, const nsol_grid *g, const double *const taps_host[3], const int32_t radius[3],
                             const void *x_dev, void *y_dev, void *tmp_dev, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    if (!x_dev || !y_dev || !taps_host || !radius) return nsol_fail(ctx, NSOL_EINVAL, "nsol_blur_sep: NULL argument");
    if (x_dev == y_dev || x_dev == tmp_dev || (tmp_dev && tmp_dev == y_dev)) return nsol_fail(ctx, NSOL_EINVAL, "nsol_blur_sep: x, y and tmp must not alias");
    NSOL_CHECK(nsol_bind_device(ctx));
    GridView gv;
    NSOL_CHECK(nsol_grid_view(ctx, g, &gv));
    for (int a = 0; a < gv.dim; ++a) {
        if (!taps_host[a]) return nsol_fail(ctx, NSOL_EINVAL, "nsol_blur_sep: taps[%d] is NULL", a);
        if (radius[a] < 0 || 2 * radius[a] + 1 > NSOL_MAX_TAPS) return nsol_fail(ctx, NSOL_EINVAL, "nsol_blur_sep: radius[%d]=%d unsupported (max %d)", a, radius[a], (NSOL_MAX_TAPS - 1) / 2);
    }
    if (gv.dtype == NSOL_F32) return blur_sep_t<float>(ctx, gv, g, taps_host, radius, x_dev, y_dev, tmp_dev, (cudaStream_t)s);
    return blur_sep_t<double>(ctx, gv, g, taps_host, radius, x_dev, y_dev, tmp_dev, (cudaStream_t)s);
}

// ---------------------------------------------------------------------------
// periodic dense convolution (arbitrary odd-sized mask, e.g. non-diagonal covariance)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void conv_dense_wrap_kernel(const T *__restrict__ in, T *__restrict__ out, const T *__restrict__ mask, long long n, long long total,
                                       int nx, int ny, int nz, int kx, int ky, int kz) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    const int cx = (kx - 1) / 2, cy = (ky - 1) / 2, cz = (kz - 1) / 2;
    for (; i < total; i += step) {
        const long long bi = i / n, r = i - bi * n;
        int ix, iy, iz;
        decode_axes<T>(r, nx, ny, ix, iy, iz);
        const T *vol = in + bi * n;
        T acc = T(0);
        for (int c = 0; c < kz; ++c) {
            int qz = (iz - (c - cz)) % nz;
            if (qz < 0) qz += nz;
            for (int b = 0; b < ky; ++b) {
                int qy = (iy - (b - cy)) % ny;
                if (qy < 0) qy += ny;
                for (int a = 0; a < kx; ++a) {
                    int qx = (ix - (a - cx)) % nx;
                    if (qx < 0) qx += nx;
                    acc += mask[((long long)c * ky + b) * kx + a] * vol[((long long)qz * ny + qy) * nx + qx];
                }
            }
        }
        out[i] = acc;
    }
}

extern "C" int nsol_conv_wrap(nsol_ctx *ctx, const nsol_grid *g, const double *kernel_host, const int64_t kshape[3], const void *x_dev,
                              void *y_dev, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    if (!kernel_host || !kshape || !x_dev || !y_dev || x_dev == y_dev) return nsol_fail(ctx, NSOL_EINVAL, "nsol_conv_wrap: bad argument");
    NSOL_CHECK(nsol_bind_device(ctx));
    GridView gv;
    NSOL_CHECK(nsol_grid_view(ctx, g, &gv));
    // kernel axes in numpy order -> (kz, ky, kx) of the kernel-axis view
    int kd[3] = {1, 1, 1};   // kz, ky, kx
    long long kn = 1;
    for (int a = 0; a < gv.dim; ++a) {
        if (kshape[a] < 1 || kshape[a] % 2 == 0 || kshape[a] > 4096) return nsol_fail(ctx, NSOL_EINVAL, "nsol_conv_wrap: mask extents must be odd and >= 1");
        kn *= kshape[a];
    }
    if (gv.dim == 1) kd[2] = (int)kshape[0];
    else if (gv.dim == 2) { kd[0] = (int)kshape[0]; kd[2] = (int)kshape[1]; }
    else { kd[0] = (int)kshape[0]; kd[1] = (int)kshape[1]; kd[2] = (int)kshape[2]; }
    cudaStream_t st = (cudaStream_t)s;
    void *mask = nullptr;
    NSOL_CUDA(ctx, cudaMallocAsync(&mask, (size_t)kn * nsol_dtype_size(gv.dtype), st));
    const long long total = gv.n * gv.batch;
    const int threads = 128;
    int rc = NSOL_OK;
    if (gv.dtype == NSOL_F32) {
        std::vector<float> h(kn);
        for (long long i = 0; i < kn; ++i) h[i] = (float)kernel_host[i];
        cudaMemcpyAsync(mask, h.data(), kn * sizeof(float), cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);   // h goes out of scope
        conv_dense_wrap_kernel<float><<<grid_1d(ctx, total, threads), threads, 0, st>>>((const float *)x_dev, (float *)y_dev, (const float *)mask,
                                                                                       gv.n, total, gv.nx, gv.ny, gv.nz, kd[2], kd[1], kd[0]);
    } else {
        cudaMemcpyAsync(mask, kernel_host, kn * sizeof(double), cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        conv_dense_wrap_kernel<double><<<grid_1d(ctx, total, threads), threads, 0, st>>>((const double *)x_dev, (double *)y_dev, (const double *)mask,
                                                                                        gv.n, total, gv.nx, gv.ny, gv.nz, kd[2], kd[1], kd[0]);
    }
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = nsol_fail(ctx, NSOL_ECUDA, "nsol_conv_wrap launch: %s", cudaGetErrorString(e));
    cudaFreeAsync(mask, st);
    return rc;
}

// ---------------------------------------------------------------------------
// stand-alone proximal maps (element-wise)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void prox_kernel(int kind, long long n, const T *__restrict__ x, const T *__restrict__ x0, T p0, T p1, T *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    const T hub = T(1) + p0 * p1;
    for (; i < n; i += step) {
        T v = x[i], r;
        if (kind == NSOL_PROX_TV_CONJ) {
            r = v / max_t(T(1), abs_t(v));
        } else if (kind == NSOL_PROX_HUBER_CONJ) {
            v = v / hub;
            r = v / max_t(T(1), abs_t(v));
        } else if (kind == NSOL_PROX_TK1_CONJ) {
            r = v / (T(1) + p0);
        } else {
            const T b = x0[i] / p1;
            if (kind == NSOL_PROX_ELL2) {
                r = (v + p0 * b) / (T(1) + p0);
            } else {
                const T d = v - b;
                const T m = max_t(abs_t(d) - p0, T(0));
                const T sgn = d > T(0) ? T(1) : (d < T(0) ? T(-1) : T(0));
                r = b + m * sgn;
            }
        }
        out[i] = r;
    }
}

extern "C" int nsol_prox_apply(nsol_ctx *ctx, int kind, int dtype, int64_t n, const void *x_dev, const void *x0_dev, double p0, double p1,
                               void *out_dev, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    if (kind < NSOL_PROX_TV_CONJ || kind > NSOL_PROX_ELL2) return nsol_fail(ctx, NSOL_EINVAL, "nsol_prox_apply: unknown kind %d", kind);
    if (n < 0 || !x_dev || !out_dev) return nsol_fail(ctx, NSOL_EINVAL, "nsol_prox_apply: bad argument");
    if (kind >= NSOL_PROX_ELL1 && (!x0_dev || p1 == 0.0)) return nsol_fail(ctx, NSOL_EINVAL, "nsol_prox_apply: ELL1/ELL2 need x0 and a non-zero x_scale");
    if (n == 0) return NSOL_OK;
    NSOL_CHECK(nsol_bind_device(ctx));
    const int threads = 256;
    const int blocks = grid_1d(ctx, n, threads);
    if (dtype == NSOL_F32)
        prox_kernel<float><<<blocks, threads, 0, (cudaStream_t)s>>>(kind, n, (const float *)x_dev, (const float *)x0_dev, (float)p0, (float)p1, (float *)out_dev);
    else if (dtype == NSOL_F64)
        prox_kernel<double><<<blocks, threads, 0, (cudaStream_t)s>>>(kind, n, (const double *)x_dev, (const double *)x0_dev, p0, p1, (double *)out_dev);
    else
        return nsol_fail(ctx, NSOL_EINVAL, "nsol_prox_apply: bad dtype");
    NSOL_LAUNCH_CHECK(ctx);
    return NSOL_OK;
}

// ---------------------------------------------------------------------------
// on-device measures: one reduction pass per iterate (no copy of the iterate to the host)
// ---------------------------------------------------------------------------
#define STATS_THREADS 256
#define STATS_MAX_BLOCKS 1024

template <int K>
__device__ __forceinline__ void stats_block_reduce(double (&v)[K], bool is_max_last, double *out_block) {
    __shared__ double sm[K][STATS_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x = v[k];
        const bool mx = is_max_last && k == K - 1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = mx ? fmax(x, y) : x + y;
        }
        if (lane == 0) sm[k][wid] = x;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const bool mx = is_max_last && k == K - 1;
            double x = lane < STATS_THREADS / 32 ? sm[k][lane] : (mx ? -INFINITY : 0.0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double y = __shfl_down_sync(0xffffffffu, x, o);
                x = mx ? fmax(x, y) : x + y;
            }
            if (lane == 0) out_block[k] = x;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(STATS_THREADS) similarity_stats_kernel(long long n, const T *__restrict__ x, double scale,
                                                                         const double *__restrict__ r, double *__restrict__ part) {
    double v[8] = {0, 0, 0, 0, 0, 0, 0, -INFINITY};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double y = (double)x[i] * scale, ri = r[i], d = y - ri;
        v[0] += y; v[1] += y * y; v[2] += ri; v[3] += ri * ri; v[4] += y * ri; v[5] += d * d; v[6] += fabs(d);
        v[7] = fmax(v[7], ri);
    }
    stats_block_reduce<8>(v, true, part + (size_t)blockIdx.x * 8);
}

template <typename T>
__global__ void __launch_bounds__(STATS_THREADS) prior_stats_kernel(GradArgs<T> a, double scale, double gamma, double *__restrict__ part) {
    double v[4] = {0, 0, 0, 0};
    const double g2 = gamma * gamma;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (long long)gridDim.x * blockDim.x) {
        int idx[3];
        decode_axes<T>(i, a.nx, a.ny, idx[0], idx[1], idx[2]);
        const double c = (double)a.in[i] * scale;
        double ss = 0.0;
        for (int k = 0; k < a.dim; ++k) {
            const int ax = (k == 0) ? 0 : ((a.dim == 3 && k == 1) ? 1 : 2);
            const double hi = (idx[ax] + 1 < a.extent[k]) ? (double)a.in[i + a.stride[k]] * scale : 0.0;
            const double d = (double)a.w[k] * hi + (-(double)a.w[k]) * c;
            ss = (k == 0) ? d * d : ss + d * d;
        }
        v[0] += sqrt(ss);
        v[1] += ss;
        v[2] += ss < g2 ? ss : 2.0 * gamma * sqrt(ss) - g2;     // loss_functions.huber (f_scale = 1)
        v[3] += c * c;
    }
    stats_block_reduce<4>(v, false, part + (size_t)blockIdx.x * 4);
}

template <int K>
__global__ void stats_final_kernel(const double *__restrict__ part, int nblocks, bool is_max_last, double *__restrict__ out) {
    double v[K];
    for (int k = 0; k < K; ++k) v[k] = (is_max_last && k == K - 1) ? -INFINITY : 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
        for (int k = 0; k < K; ++k) {
            const double x = part[(size_t)b * K + k];
            v[k] = (is_max_last && k == K - 1) ? fmax(v[k], x) : v[k] + x;
        }
    stats_block_reduce<K>(v, is_max_last, out);
}

static int stats_finish(nsol_ctx *ctx, double *part, int nblocks, int k, double *out_host, cudaStream_t st) {
    double *res = part + (size_t)STATS_MAX_BLOCKS * 8;
    if (k == 8) stats_final_kernel<8><<<1, STATS_THREADS, 0, st>>>(part, nblocks, true, res);
    else stats_final_kernel<4><<<1, STATS_THREADS, 0, st>>>(part, nblocks, false, res);
    NSOL_LAUNCH_CHECK(ctx);
    NSOL_CUDA(ctx, cudaMemcpyAsync(out_host, res, k * sizeof(double), cudaMemcpyDeviceToHost, st));
    NSOL_CUDA(ctx, cudaStreamSynchronize(st));
    NSOL_CUDA(ctx, cudaFreeAsync(part, st));
    return NSOL_OK;
}

extern "C" int nsol_similarity_stats(nsol_ctx *ctx, int dtype_x, int64_t n, const void *x_dev, double scale, const double *xref_dev,
                                     double *out_host8, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    if (n < 1 || !x_dev || !xref_dev || !out_host8) return nsol_fail(ctx, NSOL_EINVAL, "nsol_similarity_stats: bad argument");
    NSOL_CHECK(nsol_bind_device(ctx));
    cudaStream_t st = (cudaStream_t)s;
    long long want = (n + STATS_THREADS - 1) / STATS_THREADS;
    const int nb = (int)(want < STATS_MAX_BLOCKS ? want : STATS_MAX_BLOCKS);
    double *part = nullptr;
    NSOL_CUDA(ctx, cudaMallocAsync((void **)&part, sizeof(double) * (STATS_MAX_BLOCKS + 1) * 8, st));
    if (dtype_x == NSOL_F32) similarity_stats_kernel<float><<<nb, STATS_THREADS, 0, st>>>(n, (const float *)x_dev, scale, xref_dev, part);
    else if (dtype_x == NSOL_F64) similarity_stats_kernel<double><<<nb, STATS_THREADS, 0, st>>>(n, (const double *)x_dev, scale, xref_dev, part);
    else return nsol_fail(ctx, NSOL_EINVAL, "nsol_similarity_stats: bad dtype");
    NSOL_LAUNCH_CHECK(ctx);
    return stats_finish(ctx, part, nb, 8, out_host8, st);
}

extern "C" int nsol_prior_stats(nsol_ctx *ctx, const nsol_grid *g, const void *x_dev, double scale, double huber_gamma, double *out_host4,
                                nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    if (!x_dev || !out_host4 || !(huber_gamma > 0.0)) return nsol_fail(ctx, NSOL_EINVAL, "nsol_prior_stats: bad argument");
    NSOL_CHECK(nsol_bind_device(ctx));
    GridView gv;
    NSOL_CHECK(nsol_grid_view(ctx, g, &gv));
    if (gv.batch != 1) return nsol_fail(ctx, NSOL_EINVAL, "nsol_prior_stats: grid.batch must be 1");
    cudaStream_t st = (cudaStream_t)s;
    long long want = (gv.n + STATS_THREADS - 1) / STATS_THREADS;
    const int nb = (int)(want < STATS_MAX_BLOCKS ? want : STATS_MAX_BLOCKS);
    double *part = nullptr;
    NSOL_CUDA(ctx, cudaMallocAsync((void **)&part, sizeof(double) * (STATS_MAX_BLOCKS + 1) * 8, st));
    if (gv.dtype == NSOL_F32) {
        GradArgs<float> a;
        fill_grad_args<float>(gv, x_dev, nullptr, a);
        prior_stats_kernel<float><<<nb, STATS_THREADS, 0, st>>>(a, scale, huber_gamma, part);
    } else {
        GradArgs<double> a;
        fill_grad_args<double>(gv, x_dev, nullptr, a);
        prior_stats_kernel<double><<<nb, STATS_THREADS, 0, st>>>(a, scale, huber_gamma, part);
    }
    NSOL_LAUNCH_CHECK(ctx);
    return stats_finish(ctx, part, nb, 4, out_host4, st);
}
