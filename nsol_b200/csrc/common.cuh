// common.cuh -- shared host/device helpers for libnsol_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>

#include "nsol_b200.h"

struct nsol_ctx {
    int device = 0;
    int sm_count = 148;
    int64_t launches = 0;
    // tuning knobs (0 = default)
    int pd_zc = 0;
    int pd_ty = 0;
    int pd_variant = 0;
    int pd_persist_blocks = 0;   // persistent primal-dual kernel: cap on the number of CTAs (0: one per SM)
    int pd_persist = 0;     // persistent cooperative primal-dual kernel for small 2-D / 1-D problems: 0 auto, 1 whenever possible, 2 never
    // pipelined host solve (nsol_pd_plan_solve_host): the z-ordered upload / download of a large volume overlaps a wavefront of
    // iterations.  pd_pipe: 0 auto (volumes >= 64 MiB, pinned host buffers), 1 whenever possible, 2 never; pd_pipe_depth: iterations
    // the first group runs ahead of the last one (0 = 10); pd_pipe_planes: z-planes per transfer group (0 = 16)
    int pd_pipe = 0, pd_pipe_depth = 0, pd_pipe_planes = 0;
    // 2-D temporal blocking (csrc/pd_tb2d.cuh): pd_tb 0 auto (2-D problems, >= 2 iterations), 1 whenever possible, 2 never;
    // pd_tb_k iterations per pass (0 = 4); pd_tb_nr rows per thread = region height 16 / 32 / 32 for 1 / 2 / 4 (0 = chosen from the
    // problem size)
    int pd_tb = 0, pd_tb_k = 0, pd_tb_nr = 0;
    int pd_push = 0;        // linked z-slabs: 1 = boundary planes pushed by a publish kernel on a second stream (experiment), else by the boundary CTAs
    int pd_chain = 0;       // iteration chaining (per-chunk dependencies between consecutive whole-volume launches): 1 on (experiment; slower), else off
    int pd_pdl = 0;         // programmatic dependent launch of the primal-dual iteration kernels: 0 on, 2 off
    int lsmr_blocks = 0;
    int lsmr_path = 0;      // 0 auto, 1 multi-kernel (vector kernels where they apply), 2 cooperative single launch (generic phases),
                            // 3 multi-kernel with the generic kernels, 4 persistent cooperative solve built from the vector phases
    int lsmr_fuse2d = 0;    // fused 2-D forward / adjoint LSMR kernels: 0 auto (large images), 1 whenever possible, 2 never,
                            // 3 whenever possible with the first-generation kernels (lsmr_fastv.cuh; kept for comparison)
    int lsmr_tile = 0;      // 2-D persistent solve with tile-fused blur, two grid barriers per inner iteration (csrc/lsmr_tile2d.cuh): 0 on, 2 never
    int lsmr_fuse3d = 0;    // fused 3-D forward / adjoint LSMR kernels (csrc/lsmr_fused3d.cuh): same values
    int link_timeout_ms = 0; // in-kernel halo exchange: give up waiting for a neighbour after this long (0 = 5000)
    // "debug_guard" = 1: the arrays of every plan created afterwards sit between two NaN-filled guard bands; an out-of-bounds
    // read drags NaN into the result, an out-of-bounds write is found by nsol_debug_guard_check (the pool's compute-sanitizer
    // is closed -- this is the bounds check the parity tests run instead)
    int debug_guard = 0;
    std::map<void *, std::pair<void *, size_t>> guarded;   // user pointer -> (base pointer, user bytes)
    std::string err;
};

#define NSOL_GUARD_BYTES ((size_t)1 << 16)
// cudaMalloc / cudaFree of a plan array, guarded when ctx->debug_guard is set (capi.cu)
cudaError_t nsol_plan_alloc(nsol_ctx *ctx, void **ptr, size_t bytes);
void nsol_plan_free(nsol_ctx *ctx, void *ptr);

extern thread_local std::string g_nsol_create_error;

inline int nsol_fail(nsol_ctx *ctx, int code, const char *fmt, ...) __attribute__((format(printf, 3, 4)));
inline int nsol_fail(nsol_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_nsol_create_error = buf;
    return code;
}

#define NSOL_CUDA(ctx, call)                                                                    \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return nsol_fail((ctx), e__ == cudaErrorMemoryAllocation ? NSOL_ENOMEM : NSOL_ECUDA, \
                             "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

#define NSOL_CHECK(expr)          \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != NSOL_OK) return rc__; \
    } while (0)

#define NSOL_LAUNCH_CHECK(ctx)                                                                  \
    do {                                                                                        \
        (ctx)->launches++;                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return nsol_fail((ctx), NSOL_ECUDA, "%s:%d: kernel launch -> %s", __FILE__, __LINE__, \
                             cudaGetErrorString(e__));                                          \
    } while (0)

inline int nsol_bind_device(nsol_ctx *ctx) {
    int cur = -1;
    NSOL_CUDA(ctx, cudaGetDevice(&cur));
    if (cur != ctx->device) NSOL_CUDA(ctx, cudaSetDevice(ctx->device));
    return NSOL_OK;
}

inline size_t nsol_dtype_size(int dtype) { return dtype == NSOL_F32 ? 4 : 8; }

// Validated, kernel-axis view of an nsol_grid: every problem is treated as a
// (nz, ny, nx) volume.  The reference's derivative component k acts on numpy
// axis dim-1-k; the kernels use three axis "roles":
//   role X: the contiguous axis           (component 0)
//   role Y: the middle axis of a 3-D grid (component 1 when dim == 3)
//   role Z: the slowest axis              (component dim-1 when dim >= 2)
// A 2-D image (ny_img, nx) is mapped to (nz = ny_img, ny = 1, nx) so that the
// marching direction of the stencil kernels is always the slowest axis and the
// y role (shared-memory exchange between warps) only exists in 3-D.
struct GridView {
    int dim = 0;
    int nx = 1, ny = 1, nz = 1;
    long long n = 0;       // voxels per problem
    int batch = 1;
    int dtype = NSOL_F64;
    double w[3] = {0, 0, 0};   // fl(1/spacing[k]) per reference component k
    // kernel roles -> reference component index (or -1)
    int comp_x = 0, comp_y = -1, comp_z = -1;
};

inline int nsol_grid_view(nsol_ctx *ctx, const nsol_grid *g, GridView *v) {
    if (!g) return nsol_fail(ctx, NSOL_EINVAL, "grid is NULL");
    if (g->dim < 1 || g->dim > 3) return nsol_fail(ctx, NSOL_EINVAL, "grid.dim must be 1, 2 or 3 (got %d)", g->dim);
    if (g->dtype != NSOL_F64 && g->dtype != NSOL_F32) return nsol_fail(ctx, NSOL_EINVAL, "grid.dtype must be NSOL_F64 or NSOL_F32");
    if (g->batch < 1) return nsol_fail(ctx, NSOL_EINVAL, "grid.batch must be >= 1");
    long long n = 1;
    for (int a = 0; a < g->dim; ++a) {
        if (g->shape[a] < 1 || g->shape[a] > 0x7fffffffLL) return nsol_fail(ctx, NSOL_EINVAL, "grid.shape[%d] out of range", a);
        n *= g->shape[a];
        if (!(g->spacing[a] > 0.0)) return nsol_fail(ctx, NSOL_EINVAL, "grid.spacing[%d] must be > 0", a);
    }
    v->dim = g->dim;
    v->dtype = g->dtype;
    v->batch = g->batch;
    v->n = n;
    v->nx = (int)g->shape[g->dim - 1];
    v->ny = g->dim == 3 ? (int)g->shape[1] : 1;
    v->nz = g->dim >= 2 ? (int)g->shape[0] : 1;
    for (int k = 0; k < g->dim; ++k) v->w[k] = 1.0 / g->spacing[k];
    v->comp_x = 0;
    v->comp_y = g->dim == 3 ? 1 : -1;
    v->comp_z = g->dim >= 2 ? g->dim - 1 : -1;
    return NSOL_OK;
}

// primal-dual step sizes (pd_kernels.cu): rows of 8 doubles (sigma, tau, tau*lambda, theta, den_g, den_f, 0, 0)
void pd_schedule_rows(const nsol_pd_desc &d, double alpha, int iterations, double *rows);

#ifdef __CUDACC__
// ---- small aligned vector type for 128-bit global/shared accesses ----------
template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Vec {
    T v[VEC];
};

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> vec_zero() {
    Vec<T, VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = T(0);
    return r;
}

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> vec_load(const T *p) {
    return *reinterpret_cast<const Vec<T, VEC> *>(p);
}

template <typename T, int VEC>
__device__ __forceinline__ void vec_store(T *p, const Vec<T, VEC> &x) {
    *reinterpret_cast<Vec<T, VEC> *>(p) = x;
}

__device__ __forceinline__ double shfl_down_t(double x, int d) { return __shfl_down_sync(0xffffffffu, x, d); }
__device__ __forceinline__ float shfl_down_t(float x, int d) { return __shfl_down_sync(0xffffffffu, x, d); }
__device__ __forceinline__ double shfl_t(double x, int lane) { return __shfl_sync(0xffffffffu, x, lane); }
__device__ __forceinline__ float shfl_t(float x, int lane) { return __shfl_sync(0xffffffffu, x, lane); }
__device__ __forceinline__ double shfl_up_t(double x, int d) { return __shfl_up_sync(0xffffffffu, x, d); }
__device__ __forceinline__ float shfl_up_t(float x, int d) { return __shfl_up_sync(0xffffffffu, x, d); }

__device__ __forceinline__ double abs_t(double x) { return fabs(x); }
__device__ __forceinline__ float abs_t(float x) { return fabsf(x); }
__device__ __forceinline__ double max_t(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float max_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
#endif
