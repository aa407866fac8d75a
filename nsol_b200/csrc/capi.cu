// capi.cu -- context, memory/stream helpers and element-wise conversion kernels.
#include <stdlib.h>

#include "common.cuh"

thread_local std::string g_nsol_create_error;

extern "C" int nsol_version(void) { return NSOL_B200_VERSION; }

extern "C" int nsol_create(int device, nsol_ctx **out) {
    if (!out) return nsol_fail(nullptr, NSOL_EINVAL, "nsol_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return nsol_fail(nullptr, NSOL_ECUDA,
                         "nsol_create: no CUDA device (%s); libnsol_b200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0) {
        e = cudaGetDevice(&device);
        if (e != cudaSuccess) return nsol_fail(nullptr, NSOL_ECUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    }
    if (device >= count) return nsol_fail(nullptr, NSOL_EINVAL, "nsol_create: device %d out of range (%d devices)", device, count);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return nsol_fail(nullptr, NSOL_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return nsol_fail(nullptr, NSOL_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return nsol_fail(nullptr, NSOL_ECUDA, "nsol_create: device %d is sm_%d%d; this library is built for sm_100a only",
                         device, prop.major, prop.minor);
    nsol_ctx *ctx = new nsol_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    // tuning overrides from the environment (same keys as nsol_set_tuning, upper case)
    if (const char *v = getenv("NSOL_PD_VARIANT")) ctx->pd_variant = atoi(v);
    if (const char *v = getenv("NSOL_PD_TY")) ctx->pd_ty = atoi(v);
    if (const char *v = getenv("NSOL_PD_ZC")) ctx->pd_zc = atoi(v);
    if (const char *v = getenv("NSOL_LSMR_PATH")) ctx->lsmr_path = atoi(v);
    if (const char *v = getenv("NSOL_LSMR_BLOCKS")) ctx->lsmr_blocks = atoi(v);
    if (const char *v = getenv("NSOL_LSMR_FUSE2D")) ctx->lsmr_fuse2d = atoi(v);
    if (const char *v = getenv("NSOL_LSMR_FUSE3D")) ctx->lsmr_fuse3d = atoi(v);
    if (const char *v = getenv("NSOL_LSMR_TILE")) ctx->lsmr_tile = atoi(v);
    if (const char *v = getenv("NSOL_PD_PERSIST")) ctx->pd_persist = atoi(v);
    if (const char *v = getenv("NSOL_PD_PERSIST_BLOCKS")) ctx->pd_persist_blocks = atoi(v);
    if (const char *v = getenv("NSOL_PD_TB")) ctx->pd_tb = atoi(v);
    if (const char *v = getenv("NSOL_PD_TB_K")) ctx->pd_tb_k = atoi(v);
    if (const char *v = getenv("NSOL_PD_TB_NR")) ctx->pd_tb_nr = atoi(v);
    if (const char *v = getenv("NSOL_PD_PDL")) ctx->pd_pdl = atoi(v);
    if (const char *v = getenv("NSOL_PD_CHAIN")) ctx->pd_chain = atoi(v);
    if (const char *v = getenv("NSOL_PD_PUSH")) ctx->pd_push = atoi(v);
    if (const char *v = getenv("NSOL_PD_PIPE")) ctx->pd_pipe = atoi(v);
    if (const char *v = getenv("NSOL_PD_PIPE_DEPTH")) ctx->pd_pipe_depth = atoi(v);
    if (const char *v = getenv("NSOL_PD_PIPE_PLANES")) ctx->pd_pipe_planes = atoi(v);
    *out = ctx;
    return NSOL_OK;
}

extern "C" void nsol_destroy(nsol_ctx *ctx) { delete ctx; }

extern "C" const char *nsol_last_error(const nsol_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_nsol_create_error.c_str();
}

extern "C" int nsol_set_tuning(nsol_ctx *ctx, const char *key, int value) {
    if (!ctx || !key) return NSOL_EINVAL;
    if (value < 0) value = 0;
    if (!strcmp(key, "pd_zc")) ctx->pd_zc = value;
    else if (!strcmp(key, "pd_ty")) ctx->pd_ty = value;
    else if (!strcmp(key, "pd_variant")) ctx->pd_variant = value;
    else if (!strcmp(key, "lsmr_blocks")) ctx->lsmr_blocks = value;
    else if (!strcmp(key, "lsmr_path")) ctx->lsmr_path = value;
    else if (!strcmp(key, "link_timeout_ms")) ctx->link_timeout_ms = value;
    else if (!strcmp(key, "lsmr_fuse2d")) ctx->lsmr_fuse2d = value;
    else if (!strcmp(key, "lsmr_fuse3d")) ctx->lsmr_fuse3d = value;
    else if (!strcmp(key, "lsmr_tile")) ctx->lsmr_tile = value;
    else if (!strcmp(key, "debug_guard")) ctx->debug_guard = value;
    else if (!strcmp(key, "pd_persist")) ctx->pd_persist = value;
    else if (!strcmp(key, "pd_persist_blocks")) ctx->pd_persist_blocks = value;
    else if (!strcmp(key, "pd_tb")) ctx->pd_tb = value;
    else if (!strcmp(key, "pd_tb_k")) ctx->pd_tb_k = value;
    else if (!strcmp(key, "pd_tb_nr")) ctx->pd_tb_nr = value;
    else if (!strcmp(key, "pd_pdl")) ctx->pd_pdl = value;
    else if (!strcmp(key, "pd_chain")) ctx->pd_chain = value;
    else if (!strcmp(key, "pd_push")) ctx->pd_push = value;
    else if (!strcmp(key, "pd_pipe")) ctx->pd_pipe = value;
    else if (!strcmp(key, "pd_pipe_depth")) ctx->pd_pipe_depth = value;
    else if (!strcmp(key, "pd_pipe_planes")) ctx->pd_pipe_planes = value;
    else return nsol_fail(ctx, NSOL_EINVAL, "nsol_set_tuning: unknown key '%s'", key);
    return NSOL_OK;
}

// ---- guarded plan arrays (debug_guard) -------------------------------------------------------------------------
cudaError_t nsol_plan_alloc(nsol_ctx *ctx, void **ptr, size_t bytes) {
    if (!ctx->debug_guard) return cudaMalloc(ptr, bytes);
    const size_t user = (bytes + 255) / 256 * 256;
    char *base = nullptr;
    cudaError_t e = cudaMalloc((void **)&base, user + 2 * NSOL_GUARD_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaMemset(base, 0xFF, user + 2 * NSOL_GUARD_BYTES);      // all-ones words are NaN in float32 and float64
    if (e != cudaSuccess) {
        cudaFree(base);
        return e;
    }
    *ptr = base + NSOL_GUARD_BYTES;
    ctx->guarded[*ptr] = std::make_pair((void *)base, bytes);
    return cudaSuccess;
}

void nsol_plan_free(nsol_ctx *ctx, void *ptr) {
    if (!ptr) return;
    if (ctx) {
        auto it = ctx->guarded.find(ptr);
        if (it != ctx->guarded.end()) {
            cudaFree(it->second.first);
            ctx->guarded.erase(it);
            return;
        }
    }
    cudaFree(ptr);
}

__global__ void guard_check_kernel(const unsigned char *lo, size_t lo_len, const unsigned char *hi, size_t hi_len, unsigned long long *bad) {
    unsigned long long n = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = t; i < lo_len; i += stride) n += lo[i] != 0xFF;
    for (size_t i = t; i < hi_len; i += stride) n += hi[i] != 0xFF;
    if (n) atomicAdd(bad, n);
}

// Number of guard-band bytes any kernel has overwritten so far (synchronises the device).  arrays_out: guarded arrays alive.
extern "C" int nsol_debug_guard_check(nsol_ctx *ctx, int64_t *violations_out, int *arrays_out) {
    if (!ctx || !violations_out) return NSOL_EINVAL;
    NSOL_CHECK(nsol_bind_device(ctx));
    unsigned long long *bad = nullptr;
    NSOL_CUDA(ctx, cudaMalloc((void **)&bad, sizeof(unsigned long long)));
    NSOL_CUDA(ctx, cudaMemset(bad, 0, sizeof(unsigned long long)));
    NSOL_CUDA(ctx, cudaDeviceSynchronize());
    for (auto &kv : ctx->guarded) {
        const unsigned char *base = (const unsigned char *)kv.second.first;
        const size_t bytes = kv.second.second, user = (bytes + 255) / 256 * 256;
        // below: the whole band; above: from the end of the user bytes (the padding up to the 256-byte boundary counts too)
        guard_check_kernel<<<64, 256>>>(base, NSOL_GUARD_BYTES, base + NSOL_GUARD_BYTES + bytes, (user - bytes) + NSOL_GUARD_BYTES, bad);
    }
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, bad, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(bad);
    if (e != cudaSuccess) return nsol_fail(ctx, NSOL_ECUDA, "guard check: %s", cudaGetErrorString(e));
    *violations_out = (int64_t)h;
    if (arrays_out) *arrays_out = (int)ctx->guarded.size();
    return NSOL_OK;
}

extern "C" int64_t nsol_launch_count(const nsol_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int nsol_device_sm_count(const nsol_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int nsol_device_alloc(nsol_ctx *ctx, size_t bytes, void **dev) {
    if (!ctx || !dev) return NSOL_EINVAL;
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CUDA(ctx, cudaMalloc(dev, bytes ? bytes : 1));
    return NSOL_OK;
}
extern "C" int nsol_device_free(nsol_ctx *ctx, void *dev) {
    if (!ctx) return NSOL_EINVAL;
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CUDA(ctx, cudaFree(dev));
    return NSOL_OK;
}
extern "C" int nsol_host_alloc(nsol_ctx *ctx, size_t bytes, void **host) {
    if (!ctx || !host) return NSOL_EINVAL;
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CUDA(ctx, cudaHostAlloc(host, bytes ? bytes : 1, cudaHostAllocDefault));
    return NSOL_OK;
}
extern "C" int nsol_host_free(nsol_ctx *ctx, void *host) {
    if (!ctx) return NSOL_EINVAL;
    NSOL_CUDA(ctx, cudaFreeHost(host));
    return NSOL_OK;
}
extern "C" int nsol_memcpy_h2d(nsol_ctx *ctx, void *dev, const void *host, size_t bytes, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CUDA(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)s));
    return NSOL_OK;
}
extern "C" int nsol_memcpy_d2h(nsol_ctx *ctx, void *host, const void *dev, size_t bytes, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)s));
    return NSOL_OK;
}
extern "C" int nsol_memcpy_d2d(nsol_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
    return NSOL_OK;
}
extern "C" int nsol_memset_dev(nsol_ctx *ctx, void *dev, int value, size_t bytes, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CUDA(ctx, cudaMemsetAsync(dev, value, bytes, (cudaStream_t)s));
    return NSOL_OK;
}
extern "C" int nsol_stream_sync(nsol_ctx *ctx, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)s));
    return NSOL_OK;
}

// out[i] = (TO)(in[i] * f)  or  (TO)(in[i] / f), evaluated in the wider of the two types
template <typename TI, typename TO, bool DIV>
__global__ void scale_convert_kernel(long long n, const TI *in, TO *out, double f) {   // in == out is allowed
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double v = (double)in[i];
        v = DIV ? v / f : v * f;
        out[i] = (TO)v;
    }
}

template <typename TI, typename TO>
static void launch_scale_convert(nsol_ctx *ctx, long long n, const void *in, void *out, double f, int divide, cudaStream_t s) {
    int threads = 256;
    long long want = (n + threads - 1) / threads;
    int blocks = (int)(want < (long long)ctx->sm_count * 16 ? (want > 0 ? want : 1) : (long long)ctx->sm_count * 16);
    if (divide)
        scale_convert_kernel<TI, TO, true><<<blocks, threads, 0, s>>>(n, (const TI *)in, (TO *)out, f);
    else
        scale_convert_kernel<TI, TO, false><<<blocks, threads, 0, s>>>(n, (const TI *)in, (TO *)out, f);
}

// force the (lazily loaded) conversion kernels into the context -- see pd_preload in pd_kernels.cu
void nsol_preload_scale_convert() {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, scale_convert_kernel<double, double, true>);
    cudaFuncGetAttributes(&fa, scale_convert_kernel<double, double, false>);
    cudaFuncGetAttributes(&fa, scale_convert_kernel<double, float, true>);
    cudaFuncGetAttributes(&fa, scale_convert_kernel<float, double, false>);
    cudaGetLastError();
}

extern "C" int nsol_scale_convert(nsol_ctx *ctx, int64_t n, int dtype_in, const void *in_dev, int dtype_out,
                                  void *out_dev, double factor, int divide, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    if (n < 0 || !in_dev || !out_dev) return nsol_fail(ctx, NSOL_EINVAL, "nsol_scale_convert: bad arguments");
    if (n == 0) return NSOL_OK;
    NSOL_CHECK(nsol_bind_device(ctx));
    cudaStream_t st = (cudaStream_t)s;
    if (dtype_in == NSOL_F64 && dtype_out == NSOL_F64) launch_scale_convert<double, double>(ctx, n, in_dev, out_dev, factor, divide, st);
    else if (dtype_in == NSOL_F64 && dtype_out == NSOL_F32) launch_scale_convert<double, float>(ctx, n, in_dev, out_dev, factor, divide, st);
    else if (dtype_in == NSOL_F32 && dtype_out == NSOL_F64) launch_scale_convert<float, double>(ctx, n, in_dev, out_dev, factor, divide, st);
    else if (dtype_in == NSOL_F32 && dtype_out == NSOL_F32) launch_scale_convert<float, float>(ctx, n, in_dev, out_dev, factor, divide, st);
    else return nsol_fail(ctx, NSOL_EINVAL, "nsol_scale_convert: bad dtype");
    NSOL_LAUNCH_CHECK(ctx);
    return NSOL_OK;
}
