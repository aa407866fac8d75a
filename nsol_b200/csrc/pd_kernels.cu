// pd_kernels.cu -- fused Chambolle-Pock iteration (one launch per iteration) and the
// primal-dual plan behind nsol_pd_*.
//
// Replaces the body of PrimalDualSolver._run (reference nsol/primal_dual_solver.py:232-261)
// for B = grad / B_conj = grad_adj (nsol/linear_operators.py:121-169), the dual
// projections nsol/proximal_operators.py:139-140, :157-159 and the denoising
// prox maps :96-98, :118-120.
//
// Single-pass kernel.  Each thread owns VEC consecutive x-voxels of one (y, x)
// column position and marches through a chunk of z-planes:
//     p'   = proj((p + sigma * grad(xbar)) / den_g)           (dual, 3 components)
//     x'   = prox_f(x - tau * grad_adj(p'))                    (primal)
//     xbar'= x' + theta (x' - x)
// grad_adj at voxel i needs p' at i - e_k: along x it comes from the neighbouring
// lane (warp shuffle), along y from the warp below (shared memory, one
// __syncthreads per plane), along z from the thread's own previous plane
// (register).  On the low side of a tile p' is recomputed from the old p and xbar
// (one-voxel halo), so p and xbar are ping-pong buffered and x is updated in
// place.  Compulsory traffic: read xbar, p[d], x, b; write p[d], x, xbar
// = 5 + 2d words per voxel (11 words in 3-D = 88 B fp64 / 44 B fp32).
//
// Arithmetic keeps the reference's operation order and is compiled without FMA
// contraction (-fmad=false), so the float64 path is bit-identical to numpy.
#include <algorithm>
#include <vector>

#include "common.cuh"

// programmatic dependent launch (see NSOL_PD_LAUNCH): no-ops when the kernel was launched without the attribute
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename T>
struct PdArgs {
    const T *xbar_in;
    T *xbar_out;
    T *x;
    const T *b;
    const T *px_in, *py_in, *pz_in;   // kernel axis roles (see GridView)
    T *px_out, *py_out, *pz_out;
    const double *sched;              // [it][batch][8]
    const T *halo_xbar_above, *halo_xbar_below, *halo_pz_below;  // z-slab neighbours (or NULL)
    // in-kernel halo exchange over peer memory ("link", see nsol_pd_plan_link_*): all NULL when the
    // halos are refreshed by the caller (NCCL send/recv) or there is no neighbour on that side
    const unsigned *flag_below, *flag_above;   // local: generations the lower / upper neighbour has published
    unsigned *peer_flag_below, *peer_flag_above;   // the corresponding flags in the neighbours' memory
    unsigned *count_below, *count_above;       // local arrival counters of the boundary CTAs
    T *push_below_xbar;                        // lower neighbour's xbar_above halo plane (peer memory)
    T *push_above_xbar, *push_above_pz;        // upper neighbour's xbar_below / pz_below halo planes
    int *link_error;                           // set when a wait timed out
    unsigned long long link_timeout_ns;
    unsigned want, publish;                    // flag value this iteration needs / stores when its boundary is done
    int front_chunks;                          // link mode: 1 = the two boundary chunks first (they push from inside the kernel), 2 = last
    long long n;                      // voxels per problem
    long long b_stride;               // 0 (shared observation) or n
    int nx, ny, nz, zc, nchunks;
    int nsel, chunk_first, chunk_stride;   // chunks processed by this launch: chunk_first + i*chunk_stride, i < nsel
    int has_z;
    T wx, wy, wz;
    int it, batch;
    // iteration chaining (see pd_chain_begin): per-chunk completion counters of the whole-volume launches
    unsigned long long *chain_done;   // NULL: this launch does not count
    unsigned long long chain_need;    // value the counters of chunks c-1, c, c+1 must have reached before chunk c may start
    int chain;                        // 1: wait on the counters; 0: wait for the whole previous kernel (griddepcontrol.wait)
};

// ---- in-kernel z-slab halo exchange (peer memory over NVLink) --------------------------------
// Protocol (one flag per direction, living in the CONSUMER's memory):  the boundary CTAs of the
// iteration that produces state g store their boundary planes into the neighbour's halo buffer
// [g & 1], fence (system scope) and count themselves; the last one stores flag = g + 1 with
// release semantics.  The iteration that consumes state g spins on flag >= g + 1 (acquire) before it
// reads the halo.  Writing state g into buffer [g & 1] is safe once the neighbour has published
// state g - 1 (its readers of state g - 2 are then finished) -- which every writer CTA has
// observed, because it waits for exactly that flag value before it reads its own halo.
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#ifndef NSOL_LINK_INLINE
#define NSOL_LINK_INLINE __noinline__
#endif
#ifndef NSOL_LINK_TIMEOUT_MS
#define NSOL_LINK_TIMEOUT_MS 5000   // default; nsol_set_tuning("link_timeout_ms")
#endif
// every calling thread polls the flag itself (same address: one broadcast load per warp)
__device__ NSOL_LINK_INLINE void pd_link_wait(const unsigned *flag, unsigned want, int *error, unsigned long long timeout_ns) {
    if ((int)(ld_acquire_sys(flag) - want) >= 0) return;
    if (*(volatile int *)error) return;           // a neighbour already failed: do not wait again
    const unsigned long long t0 = global_timer_ns();
    while ((int)(ld_acquire_sys(flag) - want) < 0) {
        __nanosleep(64);
        if (global_timer_ns() - t0 > timeout_ns) {
            if (*(volatile int *)error == 0) {     // diagnostics: the generation that did not arrive, publish kernels started / finished by then
                ((volatile unsigned *)error)[1] = want;
                ((volatile unsigned *)error)[4] = ((volatile unsigned *)error)[2];
                ((volatile unsigned *)error)[5] = ((volatile unsigned *)error)[3];
            }
            *(volatile int *)error = 1;
            return;
        }
    }
}
// called by one thread of a boundary CTA after the CTA's peer stores are fenced
__device__ NSOL_LINK_INLINE void pd_link_signal(unsigned *count, unsigned *peer_flag, unsigned ctas, unsigned publish) {
    const unsigned done = atomicAdd(count, 1u);
    if (done + 1u == ctas) {
        atomicExch(count, 0u);
        __threadfence_system();
        st_release_sys(peer_flag, publish);
    }
}
// ---- iteration chaining ---------------------------------------------------------------------------------------------
// With programmatic dependent launch the CTAs of iteration k + 1 are resident while the last wave of iteration k drains.  Chunk c of
// iteration k + 1 reads (and overwrites the inputs of) chunks c - 1, c, c + 1 of iteration k only, so instead of waiting for the whole
// previous kernel (griddepcontrol.wait) it waits until those three chunks have counted all their CTAs: the tail of one iteration
// overlaps the head of the next.  All CTAs of iteration k have started before the first CTA of k + 1 is launched (that is when the
// launch_dependents trigger fires), so a spinning CTA never keeps a CTA it waits for off the SMs.  Release: every thread fences its
// stores, block barrier, one atomic increment; acquire: ld.acquire.gpu by one thread, block barrier (the acquire also drops the SM's
// stale L1 lines of the buffers the previous iterations read).  A counter that does not arrive within 2 s traps instead of hanging.
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <typename T>
__device__ __forceinline__ void pd_chain_begin(const PdArgs<T> &a, int chunk) {
    if (!a.chain) {
        pdl_wait();
        return;
    }
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const int c0 = chunk > 0 ? chunk - 1 : 0, c1 = chunk + 1 < a.nchunks ? chunk + 1 : a.nchunks - 1;
        for (int c = c0; c <= c1; ++c) {
            if (ld_acquire_gpu_u64(a.chain_done + c) >= a.chain_need) continue;
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_gpu_u64(a.chain_done + c) < a.chain_need) {
                __nanosleep(32);
                if (global_timer_ns() - t0 > 2000000000ull) __trap();
            }
        }
    }
    __syncthreads();
    asm volatile("fence.proxy.async;" ::: "memory");      // the bulk copies below read what the generic proxy of other SMs wrote
}
template <typename T>
__device__ __forceinline__ void pd_chain_end(const PdArgs<T> &a, int chunk) {
    if (!a.chain_done) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) atomicAdd(a.chain_done + chunk, 1ull);
}

// halo planes are written by another GPU while this kernel may already run: bypass L1
template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> vec_load_cg(const T *p) {
    Vec<T, VEC> r;
    if constexpr (sizeof(T) * VEC == 16) {
        const float4 q = __ldcg(reinterpret_cast<const float4 *>(p));
        memcpy(&r, &q, 16);
    } else if constexpr (sizeof(T) * VEC == 8) {
        const float2 q = __ldcg(reinterpret_cast<const float2 *>(p));
        memcpy(&r, &q, 8);
    } else {
        const float q = __ldcg(reinterpret_cast<const float *>(p));
        memcpy(&r, &q, 4);
    }
    return r;
}

// a*b + c with the reference's two roundings in float64 (no contraction: the float64 path is
// bit-identical to numpy); a fused multiply-add in float32 (tolerance 1e-4)
__device__ __forceinline__ double madd(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
__device__ __forceinline__ float madd(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// Division by a launch constant d.  float64: r = RN(1/d) once per thread, then q = x*r and two
// FMA-exact Newton corrections (Markstein): the result is the correctly rounded quotient, i.e.
// bit-identical to IEEE x / d, at 5 instructions instead of the ~30 of a DDIV sequence.
// float32: x * r (within 1 ulp; the float32 mode is specified to 1e-4).
template <typename T>
struct ConstDiv;
template <>
struct ConstDiv<double> {
    double d, r;
    __device__ __forceinline__ explicit ConstDiv(double d_) : d(d_), r(1.0 / d_) {}
    __device__ __forceinline__ double operator()(double x) const {
        double q = __dmul_rn(x, r);
        double e = __fma_rn(-q, d, x);
        q = __fma_rn(e, r, q);
        e = __fma_rn(-q, d, x);
        return __fma_rn(e, r, q);
    }
};
template <>
struct ConstDiv<float> {
    float r;
    __device__ __forceinline__ explicit ConstDiv(float d_) : r(1.0f / d_) {}
    __device__ __forceinline__ float operator()(float x) const { return x * r; }
};

// projection onto [-1, 1]: q / max(1, |q|) is q itself for |q| <= 1 and q/|q| = +-1 exactly
// otherwise, so the reference's division (proximal_operators.py:140,159) is a clamp, bit for bit
// float64: decided on the exponent word with integer instructions (|q| >= 1  <=>  the high word without
// its sign is >= 0x3ff00000): a handful of integer instructions instead of the ~12 of fmin(fmax()) with its
// NaN handling.  Same bits as the division for every finite q, including +-1 and -0; NaN stays NaN and
// +-Inf becomes NaN (Inf / Inf), as in numpy -- a non-finite observation must stay visible in the result.
__device__ __forceinline__ double clamp_unit(double q) {
    const int hi = __double2hiint(q);
    const unsigned mag = (unsigned)(hi & 0x7fffffff);
    const bool sat = mag >= 0x3ff00000u;
    int rhi = sat ? ((hi & (int)0x80000000) | 0x3ff00000) : hi;
    const int rlo = sat ? 0 : __double2loint(q);
    rhi = mag >= 0x7ff00000u ? 0x7ff80000 : rhi;
    return __hiloint2double(rhi, rlo);
}
__device__ __forceinline__ float clamp_unit(float q) {
    const float c = fminf(fmaxf(q, -1.0f), 1.0f);
    return fabsf(q) < __int_as_float(0x7f800000) ? c : __int_as_float(0x7fc00000);
}

// UNIT: spacing 1 (w == 1, every configuration of BASELINE.json): 1*a and (-1)*a are exact, so
// fl(fl(w*hi) + fl((-w)*lo)) == fl(hi - lo) bit for bit and the two multiplications are dropped
template <typename T, bool UNIT>
__device__ __forceinline__ T wdiff(T w, T hi, T lo) {
    return UNIT ? hi - lo : madd(w, hi, (-w) * lo);
}

template <typename T, int REG, bool UNIT = false>
__device__ __forceinline__ T dual_update(T p, T hi, T lo, T w, T sigma, const ConstDiv<T> &div_g) {
    // grad: fl(fl(w*x[i+1]) + fl((-w)*x[i]))   (scipy.ndimage.convolve order)
    T g = wdiff<T, UNIT>(w, hi, lo);
    T q = madd(sigma, g, p);                                 // primal_dual_solver.py:242-243
    if (REG != NSOL_REG_TV) q = div_g(q);                    // proximal_operators.py:158 / TK1
    if (REG != NSOL_REG_TK1) q = clamp_unit(q);              // proximal_operators.py:140,159
    return q;
}

template <typename T, int DATA>
__device__ __forceinline__ void primal_update(T xv, T bv, T div, T tau, T tl, T theta, const ConstDiv<T> &div_f, T &xn, T &xbn) {
    T yv = madd(-tau, div, xv);                              // x - tau*div  (primal_dual_solver.py:246)
    if (DATA == NSOL_DATA_L2) {
        xn = div_f(madd(tl, bv, yv));                        // (y + t b)/(1 + t)  (proximal_operators.py:120)
    } else {
        T d = yv - bv;                                       // proximal_operators.py:98
        T m = max_t(abs_t(d) - tl, T(0));
        T sgn = d > T(0) ? T(1) : (d < T(0) ? T(-1) : T(0));
        xn = madd(m, sgn, bv);
    }
    xbn = madd(theta, xn - xv, xn);                          // primal_dual_solver.py:253
}

// Link mode, start of a CTA: the bottom chunk reads the lower neighbour's planes, the top chunk
// the upper neighbour's plane; wait until they have been published.  (The boundary chunks are
// scheduled first and the flags were raised early in the neighbours' previous iteration, so this
// normally falls straight through.)  Every thread polls for itself.
template <typename T>
__device__ __forceinline__ void pd_link_begin(const PdArgs<T> &a, int z0, int z1) {
    if (z0 == 0 && a.flag_below) pd_link_wait(a.flag_below, a.want, a.link_error, a.link_timeout_ns);
    if (z1 == a.nz && a.flag_above) pd_link_wait(a.flag_above, a.want, a.link_error, a.link_timeout_ns);
}

// Link mode, end of a CTA: boundary CTAs copy the boundary planes of the NEW state (their own
// stores of this launch) into the neighbours' receive slots, fence (system scope) and count
// themselves; the last CTA of a side raises the neighbour's flag.  hoff = this thread's element
// offset inside a plane.  Block-uniform control flow.
template <typename T, int VEC>
__device__ __forceinline__ void pd_link_finish(const PdArgs<T> &a, int z0, int z1, long long hoff, bool active) {
    const bool bot = a.push_below_xbar && z0 == 0;
    const bool top = a.push_above_xbar && z1 == a.nz;
    if (!bot && !top) return;
    if (active) {
        if (bot) vec_store<T, VEC>(a.push_below_xbar + hoff, vec_load<T, VEC>(a.xbar_out + hoff));
        if (top) {
            const long long o = (long long)(a.nz - 1) * a.nx * a.ny + hoff;
            vec_store<T, VEC>(a.push_above_xbar + hoff, vec_load<T, VEC>(a.xbar_out + o));
            vec_store<T, VEC>(a.push_above_pz + hoff, vec_load<T, VEC>(a.pz_out + o));
        }
    }
    // one system-scope fence per CTA, by the thread that signals, after the block barrier (the pattern of a cooperative-groups grid
    // barrier: the fence is cumulative over the writes the barrier orders before it) -- not one per thread
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence_system();
        const unsigned ctas = gridDim.x * gridDim.y;
        if (bot) pd_link_signal(a.count_below, a.peer_flag_below, ctas, a.publish);
        if (top) pd_link_signal(a.count_above, a.peer_flag_above, ctas, a.publish);
    }
}

// everything one thread loads for plane z: p, x, b at z; xbar at z+1 (own voxels and, if the CTA
// also processes plane z+1, its x/y halos); the low-side halos of p needed to recompute p'
template <typename T, int VEC>
struct PdStep {
    Vec<T, VEC> px, py, pz, x, b, xbn, hup, hdn, pydn;
    T pxl, xr, xl;
};

// register budget: the float32 kernel fits 128 registers without spills (2 x 256 threads per SM
// resident), the float64 kernel needs ~160 for its two in-flight planes of double2 loads
#ifndef NSOL_PD_MINB_F64
#define NSOL_PD_MINB_F64 1
#endif
// 2-D / 1-D float64 (128-thread CTAs, no shared memory): the kernel is latency-bound at 104 registers = 16 warps per
// SM (ncu: issue slots 26 %, long-scoreboard 10); capping the registers buys warps
#ifndef NSOL_PD_MINB_F64_2D
#define NSOL_PD_MINB_F64_2D 3   // 80 registers (60 B of spills): 6 CTAs = 24 warps per SM; batched 1024^2 sweep 0.906 -> 0.945 of the roofline
#endif
#ifndef NSOL_PD_MINB_F32
#define NSOL_PD_MINB_F32 2
#endif
// One CTA's share of an iteration: tile (bx, by), z-chunk / batch member bzi.  Called by pd_iter_kernel with its block
// index and by the persistent cooperative kernel (small problems) with virtual block indices.
template <typename T, int VEC, bool HAS_Y, int REG, int DATA, bool LINK, bool UNIT>
__device__ __forceinline__ void pd_iter_body(const PdArgs<T> &a, const unsigned bx, const unsigned by, const unsigned bzi) {
    using V = Vec<T, VEC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int TY = HAS_Y ? (int)blockDim.y : 1;
    const int ty = HAS_Y ? (int)threadIdx.y : 0;
    const int tile_w = (int)blockDim.x * VEC;
    const int x0 = (int)bx * tile_w + (int)threadIdx.x * VEC;
    const int y0 = HAS_Y ? (int)by * TY : 0;
    const int y = y0 + ty;
    int chunk = a.chunk_first + (int)(bzi % (unsigned)a.nsel) * a.chunk_stride;
    if (LINK && a.front_chunks == 1) chunk = chunk == 0 ? 0 : (chunk == 1 ? a.nchunks - 1 : chunk - 1);
    if (LINK && a.front_chunks == 2) chunk = chunk + 2 < a.nchunks ? chunk + 1 : (chunk + 2 == a.nchunks ? 0 : a.nchunks - 1);
    const int bz = (int)(bzi / (unsigned)a.nsel);
    const int z0 = chunk * a.zc;
    const int z1 = min(a.nz, z0 + a.zc);
    const bool xin = x0 < a.nx;
    const bool active = xin && (y < a.ny);
    const bool has_z = a.has_z != 0;

    const double *srow = a.sched + ((long long)a.it * a.batch + bz) * 8;
    const T sigma = (T)srow[0], tau = (T)srow[1], tl = (T)srow[2], theta = (T)srow[3];
    const ConstDiv<T> div_g((T)srow[4]), div_f((T)srow[5]);
    const T wx = a.wx, wy = a.wy, wz = a.wz;

    const long long sz = (long long)a.nx * a.ny;
    const long long row = (long long)y * a.nx + x0;               // offset inside a plane
    const long long hrow = (long long)bz * sz + row;               // offset inside a halo plane array
    // running element offsets of plane z (advanced by sz per plane: no 64-bit multiplies in the loop)
    long long off = (long long)bz * a.n + (long long)z0 * sz + row;   // into x, xbar, p
    long long boff = (long long)bz * a.b_stride + (long long)z0 * sz + row;   // into b

    // shared tiles (3-D only): xbar rows [2][TY+2][tile_w], p'_y rows [2][TY+1][tile_w]
    T *s_xb = reinterpret_cast<T *>(smem_raw);
    T *s_py = s_xb + 2 * (TY + 2) * tile_w;
    const int s_col = (int)threadIdx.x * VEC;

    // edge roles
    const bool need_r = active && lane == 31 && (x0 + VEC < a.nx);   // right x-halo
    const bool need_l = active && lane == 0 && (x0 > 0);             // left x-halo
    const bool up_warp = HAS_Y && ty == TY - 1;
    const bool dn_warp = HAS_Y && ty == 0;
    const bool need_up = up_warp && xin && (y + 1 < a.ny);
    const bool need_dn = dn_warp && active && (y > 0);

    // own voxels of xbar plane zq located at element offset o; planes -1 / nz come from the z-slab
    // halos of the neighbouring rank (or are the zero boundary)
    auto load_xbar_own = [&](int zq, long long o) -> V {
        if (zq < 0 || zq >= a.nz) {
            const T *ptr = zq < 0 ? a.halo_xbar_below : a.halo_xbar_above;
            return (active && ptr) ? vec_load_cg<T, VEC>(ptr + hrow) : vec_zero<T, VEC>();
        }
        return active ? vec_load<T, VEC>(a.xbar_in + o) : vec_zero<T, VEC>();
    };
    // loads of plane zq whose offsets are o (x, xbar, p) and bo (b).  Only xbn must be zero in
    // inactive threads (it supplies the Dirichlet zero to the neighbours); the other fields of an
    // inactive thread are never stored.
    auto load_step = [&](int zq, long long o, long long bo, PdStep<T, VEC> &s) {
        s.xbn = has_z ? load_xbar_own(zq + 1, o + sz) : vec_zero<T, VEC>();
        if (active) {
            s.px = vec_load<T, VEC>(a.px_in + o);
            if (HAS_Y) s.py = vec_load<T, VEC>(a.py_in + o);
            if (has_z) s.pz = vec_load<T, VEC>(a.pz_in + o);
            s.x = vec_load<T, VEC>(a.x + o);
            s.b = vec_load<T, VEC>(a.b + bo);
        }
        s.pxl = need_l ? a.px_in[o - 1] : T(0);
        if (need_dn) s.pydn = vec_load<T, VEC>(a.py_in + o - a.nx);
        const bool nx_plane = zq + 1 < z1;     // halos of the next plane (only if this CTA processes it)
        s.xr = (need_r && nx_plane) ? a.xbar_in[o + sz + VEC] : T(0);
        s.xl = (need_l && nx_plane) ? a.xbar_in[o + sz - 1] : T(0);
        s.hup = (need_up && nx_plane) ? vec_load<T, VEC>(a.xbar_in + o + sz + a.nx) : vec_zero<T, VEC>();
        s.hdn = (need_dn && nx_plane) ? vec_load<T, VEC>(a.xbar_in + o + sz - a.nx) : vec_zero<T, VEC>();
    };

    // ---- prologue: plane z0 --------------------------------------------------
    pd_chain_begin(a, chunk);       // everything above touched only launch constants and the step-size table
    if (LINK) pd_link_begin(a, z0, z1);
    PdStep<T, VEC> cur;
    load_step(z0, off, boff, cur);
    V xb_c = load_xbar_own(z0, off);
    T xr_c = need_r ? a.xbar_in[off + VEC] : T(0);
    T xl_c = need_l ? a.xbar_in[off - 1] : T(0);
    V pz_prev = vec_zero<T, VEC>();
    if (has_z && active && (z0 > 0 || a.halo_pz_below)) {
        V xb_m = load_xbar_own(z0 - 1, off - sz);
        V pz_m = z0 > 0 ? vec_load<T, VEC>(a.pz_in + off - sz) : vec_load_cg<T, VEC>(a.halo_pz_below + hrow);
#pragma unroll
        for (int v = 0; v < VEC; ++v) pz_prev.v[v] = dual_update<T, REG, UNIT>(pz_m.v[v], xb_c.v[v], xb_m.v[v], wz, sigma, div_g);
    }
    if (HAS_Y) {
        T *buf = s_xb + (z0 & 1) * (TY + 2) * tile_w;
        vec_store<T, VEC>(buf + (ty + 1) * tile_w + s_col, xb_c);
        if (up_warp) {
            V h = need_up ? vec_load<T, VEC>(a.xbar_in + off + a.nx) : vec_zero<T, VEC>();
            vec_store<T, VEC>(buf + (TY + 1) * tile_w + s_col, h);
        }
        if (dn_warp) {
            V h = need_dn ? vec_load<T, VEC>(a.xbar_in + off - a.nx) : vec_zero<T, VEC>();
            vec_store<T, VEC>(buf + s_col, h);
        }
        __syncthreads();
    }

    // ---- march through the chunk (software pipelined: loads of plane z+1 are in flight
    //      while plane z is computed) ------------------------------------------------------
    // One plane: `cur` holds the loads of plane z, `nxt` receives those of plane z+1.  The loop
    // below calls it with the two register sets swapped on alternate planes, so no state is copied.
    auto plane_step = [&](int z, PdStep<T, VEC> &cur, PdStep<T, VEC> &nxt) {
        const bool more = (z + 1 < z1);   // plane z+1 is processed by this CTA
        if (more) load_step(z + 1, off + sz, boff + sz, nxt);

        // neighbours of the current plane
        V xup = vec_zero<T, VEC>(), xdn = vec_zero<T, VEC>();
        T *xbuf_c = s_xb + (z & 1) * (TY + 2) * tile_w;
        T *xbuf_n = s_xb + ((z + 1) & 1) * (TY + 2) * tile_w;
        T *pbuf = s_py + (z & 1) * (TY + 1) * tile_w;
        if (HAS_Y) {
            xup = vec_load<T, VEC>(xbuf_c + (ty + 2) * tile_w + s_col);
            if (dn_warp) xdn = vec_load<T, VEC>(xbuf_c + s_col);
        }
        T x_right = shfl_down_t(xb_c.v[0], 1);
        if (lane == 31) x_right = xr_c;

        // dual update
        V pnx, pny = vec_zero<T, VEC>(), pnz = vec_zero<T, VEC>();
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            T hi = (v + 1 < VEC) ? xb_c.v[(v + 1) % VEC] : x_right;
            pnx.v[v] = dual_update<T, REG, UNIT>(cur.px.v[v], hi, xb_c.v[v], wx, sigma, div_g);
            if (HAS_Y) pny.v[v] = dual_update<T, REG, UNIT>(cur.py.v[v], xup.v[v], xb_c.v[v], wy, sigma, div_g);
            if (has_z) pnz.v[v] = dual_update<T, REG, UNIT>(cur.pz.v[v], cur.xbn.v[v], xb_c.v[v], wz, sigma, div_g);
        }
        if (active) {
            vec_store<T, VEC>(a.px_out + off, pnx);
            if (HAS_Y) vec_store<T, VEC>(a.py_out + off, pny);
            if (has_z) vec_store<T, VEC>(a.pz_out + off, pnz);
        }
        // p'_x of the voxel left of this thread's first voxel
        T pnx_left = shfl_up_t(pnx.v[VEC - 1], 1);
        if (lane == 0) pnx_left = need_l ? dual_update<T, REG, UNIT>(cur.pxl, xb_c.v[0], xl_c, wx, sigma, div_g) : T(0);

        V pny_dn = vec_zero<T, VEC>();
        if (HAS_Y) {
            // publish p'_y of this row (and of the halo row below the tile), and the next xbar plane
            vec_store<T, VEC>(pbuf + (ty + 1) * tile_w + s_col, pny);
            if (dn_warp) {
                V h = vec_zero<T, VEC>();
                if (need_dn) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) h.v[v] = dual_update<T, REG, UNIT>(cur.pydn.v[v], xb_c.v[v], xdn.v[v], wy, sigma, div_g);
                }
                vec_store<T, VEC>(pbuf + s_col, h);
            }
            if (more) {
                vec_store<T, VEC>(xbuf_n + (ty + 1) * tile_w + s_col, cur.xbn);
                if (up_warp) vec_store<T, VEC>(xbuf_n + (TY + 1) * tile_w + s_col, cur.hup);
                if (dn_warp) vec_store<T, VEC>(xbuf_n + s_col, cur.hdn);
            }
            __syncthreads();
            pny_dn = vec_load<T, VEC>(pbuf + ty * tile_w + s_col);
        }

        // primal update + over-relaxation
        V xnew, xbnew;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            T lo = (v == 0) ? pnx_left : pnx.v[(v + VEC - 1) % VEC];
            T div = wdiff<T, UNIT>(wx, lo, pnx.v[v]);                                   // Dx^T p_x
            if (HAS_Y) div = div + wdiff<T, UNIT>(wy, pny_dn.v[v], pny.v[v]);            // += Dy^T p_y
            if (has_z) div = div + wdiff<T, UNIT>(wz, pz_prev.v[v], pnz.v[v]);           // += Dz^T p_z
            primal_update<T, DATA>(cur.x.v[v], cur.b.v[v], div, tau, tl, theta, div_f, xnew.v[v], xbnew.v[v]);
        }
        if (active) {
            vec_store<T, VEC>(a.x + off, xnew);
            vec_store<T, VEC>(a.xbar_out + off, xbnew);
        }

        xb_c = cur.xbn;
        pz_prev = pnz;
        xr_c = cur.xr;
        xl_c = cur.xl;
        off += sz;
        boff += sz;
    };
    PdStep<T, VEC> alt;
    for (int z = z0; z < z1; z += 2) {
        plane_step(z, cur, alt);
        if (z + 1 < z1) plane_step(z + 1, alt, cur);
    }
    if (LINK) pd_link_finish<T, VEC>(a, z0, z1, hrow, active);
    pd_chain_end(a, chunk);
}

template <typename T, int VEC, bool HAS_Y, int REG, int DATA, bool LINK, bool UNIT>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? NSOL_PD_MINB_F32 : (HAS_Y ? NSOL_PD_MINB_F64 : NSOL_PD_MINB_F64_2D)) pd_iter_kernel(const PdArgs<T> a) {
    pdl_launch_dependents();
    pd_iter_body<T, VEC, HAS_Y, REG, DATA, LINK, UNIT>(a, blockIdx.x, blockIdx.y, blockIdx.z);
}

// ---- persistent variant for small 2-D / 1-D problems (BASELINE configs 1 and 2: one 256^2 / 1024^2 image) --------------
// The whole state (x, xbar, b, p: 2.5 MB at 256^2, 42 MB at 1024^2 in float64) lives in L2, so an iteration costs a few
// microseconds of work and as much again in launch latency.  Here ONE cooperative launch runs all n iterations: every CTA
// loops over its share of the virtual blocks of an iteration (same tiles, same arithmetic as pd_iter_kernel -- bit-identical),
// then the grid synchronises and the ping-pong buffers swap.  Step sizes come from the same device table.
#include <cooperative_groups.h>
template <typename T, int VEC, int REG, int DATA, bool UNIT>
__global__ void __launch_bounds__(128) pd_iter_persist_kernel(PdArgs<T> a, int iterations, unsigned vgx, unsigned vgz) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const unsigned total = vgx * vgz;
    for (int i = 0; i < iterations; ++i) {
        for (unsigned vb = blockIdx.x; vb < total; vb += gridDim.x) pd_iter_body<T, VEC, false, REG, DATA, false, UNIT>(a, vb % vgx, 0u, vb / vgx);
        // the state this iteration wrote becomes the input of the next one
        const T *xb = a.xbar_in;
        a.xbar_in = a.xbar_out;
        a.xbar_out = const_cast<T *>(xb);
        const T *t0 = a.px_in;
        a.px_in = a.px_out;
        a.px_out = const_cast<T *>(t0);
        const T *t1 = a.pz_in;
        a.pz_in = a.pz_out;
        a.pz_out = const_cast<T *>(t1);
        a.it += 1;
        if (i + 1 < iterations) grid.sync();
    }
}

#include "pd_bulk_kernel.cuh"
#include "pd_tb2d.cuh"

#ifndef NSOL_PD_DEFAULT_VARIANT_F64
#define NSOL_PD_DEFAULT_VARIANT_F64 2
#endif
#ifndef NSOL_PD_DEFAULT_VARIANT_F32
#define NSOL_PD_DEFAULT_VARIANT_F32 1
#endif

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
struct nsol_pd_plan {
    nsol_ctx *ctx = nullptr;
    nsol_pd_desc desc;
    GridView gv;
    std::vector<double> alpha;
    size_t esz = 8;
    // device state
    void *x = nullptr, *b = nullptr;
    void *x_alt = nullptr;        // second x array of the 2-D temporal-blocking kernel (x ping-pongs between passes there)
    void *xbar[2] = {nullptr, nullptr};
    void *p[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    void *stage = nullptr;        // float64 staging for host transfers
    size_t stage_bytes = 0;
    double *sched = nullptr;      // device schedule table
    int sched_cap = 0;            // iterations covered by the table
    int cur = 0;                  // ping-pong index holding the current xbar/p
    int it = 0;
    bool ready = false;
    size_t bytes = 0;
    const void *halo_above = nullptr, *halo_below = nullptr, *halo_pz_below = nullptr;
    // in-kernel halo exchange over peer memory (nsol_pd_plan_link_*)
    char *link_block = nullptr;       // this rank's block: flags, counters, receive buffers
    size_t link_plane = 0;            // bytes of one halo plane slot
    char *peer_below = nullptr, *peer_above = nullptr;   // the neighbours' blocks mapped into this process
    bool peer_below_ipc = false, peer_above_ipc = false;
    bool link_on = false;
    bool link_fresh = false;          // the neighbours hold the halos of the current state
    unsigned link_pub = 0;            // generations published so far (= index of the next one)
    // pipelined host solve (nsol_pd_plan_solve_host): copy streams and one event per transfer group
    cudaStream_t pipe_up = nullptr, pipe_dn = nullptr;
    cudaEvent_t pipe_fork = nullptr;
    std::vector<cudaEvent_t> pipe_ev_up, pipe_ev_x, pipe_ev_dn;
    int pipe_groups_last = 0, pipe_depth_last = 0;   // what the last solve did (0 groups: plain sequence)
    // link mode, whole-slab iterations: the boundary planes are pushed to the neighbours by a small publish kernel on a second
    // stream after each iteration kernel instead of by the boundary CTAs themselves (pd_launch_iteration)
    cudaStream_t push_stream = nullptr;
    cudaEvent_t push_ev_iter[8] = {nullptr}, push_ev_done[8] = {nullptr};   // rings: an event is never re-recorded while a wait on it may be pending
    unsigned long long push_count = 0;          // publish kernels queued on push_stream so far
    unsigned long long *chain_done = nullptr;   // iteration chaining: per-chunk completion counters
    int chain_chunks = 0, chain_zc = 0;         // geometry the counters belong to
    unsigned long long chain_gen = 0;           // whole-volume launches counted since the counters were cleared
    bool chain_valid = false;                   // false: something else has touched the state since (reset, chunk-range launch, ...)
    int pipe_dir = 1;                 // +1: groups travel bottom-up, -1: top-down (odd ranks of a linked z-slab decomposition)
    unsigned pipe_g0 = 0;             // link generation of the start state of the running pipelined solve
};

static int pd_push_join(nsol_pd_plan *pl, cudaStream_t s);
static int pd_push_setup(nsol_pd_plan *pl);

// layout of a link block (all offsets 256-byte aligned)
enum { LINK_FLAG_BELOW = 0, LINK_FLAG_ABOVE = 64, LINK_COUNT_BELOW = 128, LINK_COUNT_ABOVE = 192, LINK_ERROR = 224 /* + 5 diagnostic words */, LINK_COUNT_PUB = 252,
       LINK_HEADER = 256 };
// receive slots after the header: [parity][0 = xbar_above, 1 = xbar_below, 2 = pz_below]
static inline size_t link_slot(const nsol_pd_plan *pl, int parity, int which) {
    return LINK_HEADER + (size_t)(parity * 3 + which) * pl->link_plane;
}

// step sizes: reference nsol/primal_dual_solver.py:278-283, :302-306 (ALG2), :321-337, :356-358
// (ALG3), :374-379, :398-403 (AHMOD); float64 host arithmetic in the reference's order.
void pd_schedule_rows(const nsol_pd_desc &d, double alpha, int iterations, double *rows /* [iterations][8] */) {
    const double L2 = d.L2;
    const double lmbda = 1.0 / alpha;
    double tau, sigma, gamma;
    if (d.alg == NSOL_ALG3) {
        double g = lmbda, delta = 0.05;
        double mu = 2.0 * sqrt(g * delta / L2);
        gamma = 1.0 / (1.0 + mu);       // theta, carried in the gamma slot
        sigma = mu / (2.0 * delta);
        tau = mu / (2.0 * g);
    } else if (d.alg == NSOL_ALG2_AHMOD) {
        tau = 0.02;
        sigma = 4.0 / (L2 * tau);
        gamma = 0.35 * lmbda;
    } else {
        tau = 1.0 / sqrt(L2);
        sigma = 1.0 / (L2 * tau);
        gamma = 0.35 * lmbda;
    }
    for (int i = 0; i < iterations; ++i) {
        double *r = rows + (size_t)i * 8;
        double theta;
        r[0] = sigma;
        r[1] = tau;
        r[2] = tau * lmbda;
        if (d.alg == NSOL_ALG3) {
            theta = gamma;
        } else {
            theta = 1.0 / sqrt(1.0 + 2.0 * gamma * tau);
            tau = tau * theta;
            sigma = sigma / theta;
            if (d.alg == NSOL_ALG2_AHMOD) theta = 0.0;
        }
        r[3] = theta;
        r[4] = d.reg == NSOL_REG_HUBER ? 1.0 + r[0] * d.huber_gamma : (d.reg == NSOL_REG_TK1 ? 1.0 + r[0] : 1.0);
        r[5] = 1.0 + r[2];
        r[6] = 0.0;
        r[7] = 0.0;
    }
}

static int pd_ensure_schedule(nsol_pd_plan *pl, int upto) {
    if (upto <= pl->sched_cap) return NSOL_OK;
    nsol_ctx *ctx = pl->ctx;
    int cap = pl->sched_cap ? pl->sched_cap : 128;
    while (cap < upto) cap *= 2;
    const int batch = pl->gv.batch;
    std::vector<double> rows((size_t)cap * batch * 8);
    std::vector<double> one((size_t)cap * 8);
    for (int bi = 0; bi < batch; ++bi) {
        pd_schedule_rows(pl->desc, pl->alpha[bi], cap, one.data());
        for (int i = 0; i < cap; ++i) memcpy(&rows[((size_t)i * batch + bi) * 8], &one[(size_t)i * 8], 8 * sizeof(double));
    }
    double *dev = nullptr;
    NSOL_CUDA(ctx, cudaMalloc(&dev, rows.size() * sizeof(double)));
    // synchronous copy: the old table may still be read by queued launches, so it is
    // released only after the device is idle
    cudaError_t e = cudaMemcpy(dev, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(dev);
        return nsol_fail(ctx, NSOL_ECUDA, "schedule upload: %s", cudaGetErrorString(e));
    }
    if (pl->sched) {
        cudaDeviceSynchronize();
        cudaFree(pl->sched);
    }
    pl->sched = dev;
    pl->sched_cap = cap;
    return NSOL_OK;
}

extern "C" void nsol_pd_plan_destroy(nsol_pd_plan *pl) {
    if (!pl) return;
    if (pl->ctx) nsol_bind_device(pl->ctx);
    nsol_plan_free(pl->ctx, pl->x);
    nsol_plan_free(pl->ctx, pl->x_alt);
    nsol_plan_free(pl->ctx, pl->b);
    for (int i = 0; i < 2; ++i) {
        nsol_plan_free(pl->ctx, pl->xbar[i]);
        for (int k = 0; k < 3; ++k) nsol_plan_free(pl->ctx, pl->p[i][k]);
    }
    if (pl->push_stream) {
        cudaStreamSynchronize(pl->push_stream);
        cudaStreamDestroy(pl->push_stream);
        for (int i = 0; i < 8; ++i) {
            cudaEventDestroy(pl->push_ev_iter[i]);
            cudaEventDestroy(pl->push_ev_done[i]);
        }
    }
    cudaFree(pl->stage);
    cudaFree(pl->sched);
    cudaFree(pl->chain_done);
    for (auto *v : {&pl->pipe_ev_up, &pl->pipe_ev_x, &pl->pipe_ev_dn})
        for (cudaEvent_t e : *v) cudaEventDestroy(e);
    if (pl->pipe_fork) cudaEventDestroy(pl->pipe_fork);
    if (pl->pipe_up) cudaStreamDestroy(pl->pipe_up);
    if (pl->pipe_dn) cudaStreamDestroy(pl->pipe_dn);
    if (pl->peer_below && pl->peer_below_ipc) cudaIpcCloseMemHandle(pl->peer_below);
    if (pl->peer_above && pl->peer_above_ipc) cudaIpcCloseMemHandle(pl->peer_above);
    cudaFree(pl->link_block);
    delete pl;
}

static int pd_validate_desc(nsol_ctx *ctx, const nsol_pd_desc *desc, const GridView &gv);

extern "C" int nsol_pd_plan_create(nsol_ctx *ctx, const nsol_pd_desc *desc, nsol_pd_plan **out) {
    if (!ctx) return NSOL_EINVAL;
    if (!desc || !out) return nsol_fail(ctx, NSOL_EINVAL, "nsol_pd_plan_create: NULL argument");
    *out = nullptr;
    NSOL_CHECK(nsol_bind_device(ctx));
    GridView gv;
    NSOL_CHECK(nsol_grid_view(ctx, &desc->grid, &gv));
    NSOL_CHECK(pd_validate_desc(ctx, desc, gv));

    nsol_pd_plan *pl = new nsol_pd_plan();
    pl->ctx = ctx;
    pl->desc = *desc;
    pl->gv = gv;
    pl->alpha.assign(desc->alpha, desc->alpha + gv.batch);
    pl->desc.alpha = nullptr;
    pl->esz = nsol_dtype_size(gv.dtype);
    const size_t vol_bytes = (size_t)gv.n * gv.batch * pl->esz;
    const size_t b_bytes = (size_t)gv.n * (desc->b_batched ? gv.batch : 1) * pl->esz;
    auto alloc = [&](void **ptr, size_t bytes) -> bool {
        cudaError_t e = nsol_plan_alloc(ctx, ptr, bytes);
        if (e != cudaSuccess) {
            nsol_fail(ctx, NSOL_ENOMEM, "pd plan: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
            return false;
        }
        pl->bytes += bytes;
        return true;
    };
    bool ok = alloc(&pl->x, vol_bytes) && alloc(&pl->b, b_bytes) && alloc(&pl->xbar[0], vol_bytes) && alloc(&pl->xbar[1], vol_bytes);
    for (int i = 0; ok && i < 2; ++i)
        for (int k = 0; ok && k < gv.dim; ++k) ok = alloc(&pl->p[i][k], vol_bytes);
    if (!ok) {
        nsol_pd_plan_destroy(pl);
        return NSOL_ENOMEM;
    }
    int rc = pd_ensure_schedule(pl, 128);
    if (rc != NSOL_OK) {
        nsol_pd_plan_destroy(pl);
        return rc;
    }
    *out = pl;
    return NSOL_OK;
}

static int pd_validate_desc(nsol_ctx *ctx, const nsol_pd_desc *desc, const GridView &gv) {
    if (desc->reg < NSOL_REG_TV || desc->reg > NSOL_REG_TK1) return nsol_fail(ctx, NSOL_EINVAL, "pd: unknown regulariser %d", desc->reg);
    if (desc->data != NSOL_DATA_L1 && desc->data != NSOL_DATA_L2) return nsol_fail(ctx, NSOL_EINVAL, "pd: unknown data term %d", desc->data);
    if (desc->alg < NSOL_ALG2 || desc->alg > NSOL_ALG3) return nsol_fail(ctx, NSOL_EINVAL, "pd: unknown alg_type %d", desc->alg);
    if (!(desc->L2 > 0.0)) return nsol_fail(ctx, NSOL_EINVAL, "pd: L2 must be > 0");
    if (desc->x_scale == 0.0 || desc->x0_scale == 0.0 || desc->b_scale == 0.0)
        return nsol_fail(ctx, NSOL_EINVAL, "pd: x_scale, x0_scale and b_scale must be non-zero");
    if (!desc->alpha) return nsol_fail(ctx, NSOL_EINVAL, "pd: alpha is NULL");
    for (int i = 0; i < gv.batch; ++i)
        if (!(desc->alpha[i] > 0.0)) return nsol_fail(ctx, NSOL_EINVAL, "pd: alpha[%d] must be > 0", i);
    return NSOL_OK;
}

extern "C" int nsol_pd_plan_update(nsol_pd_plan *pl, const nsol_pd_desc *desc) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!desc) return nsol_fail(ctx, NSOL_EINVAL, "pd update: desc is NULL");
    NSOL_CHECK(nsol_bind_device(ctx));
    GridView gv;
    NSOL_CHECK(nsol_grid_view(ctx, &desc->grid, &gv));
    if (gv.dim != pl->gv.dim || gv.nx != pl->gv.nx || gv.ny != pl->gv.ny || gv.nz != pl->gv.nz || gv.batch != pl->gv.batch ||
        gv.dtype != pl->gv.dtype || (desc->b_batched != 0) != (pl->desc.b_batched != 0))
        return nsol_fail(ctx, NSOL_EINVAL, "pd update: grid, dtype and batch of a plan cannot change");
    NSOL_CHECK(pd_validate_desc(ctx, desc, gv));
    pl->gv = gv;   // spacing may change
    pl->desc = *desc;
    pl->alpha.assign(desc->alpha, desc->alpha + gv.batch);
    pl->desc.alpha = nullptr;
    pl->sched_cap = 0;      // step-size table is rebuilt by the next iterate
    pl->chain_valid = false;
    pl->ready = false;
    pl->it = 0;
    return pd_ensure_schedule(pl, 128);
}

extern "C" size_t nsol_pd_plan_bytes(const nsol_pd_plan *pl) { return pl ? pl->bytes + pl->stage_bytes : 0; }
extern "C" int nsol_pd_plan_iterations_done(const nsol_pd_plan *pl) { return pl ? pl->it : 0; }

static int pd_ensure_stage(nsol_pd_plan *pl, size_t bytes) {
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

// x = xbar = x0 / x0_scale, b' = b / b_scale (nsol/solver.py:35-41, proximal_operators.py:97,119), p = 0
static int pd_reset_common(nsol_pd_plan *pl, int src_dtype, const void *b_src, const void *x0_src, cudaStream_t s) {
    nsol_ctx *ctx = pl->ctx;
    NSOL_CHECK(pd_push_join(pl, s));            // a queued publish kernel may still read the state this reset overwrites
    const GridView &gv = pl->gv;
    const long long nb = gv.n * (pl->desc.b_batched ? gv.batch : 1);
    const long long nv = gv.n * gv.batch;
    NSOL_CHECK(nsol_scale_convert(ctx, nb, src_dtype, b_src, gv.dtype, pl->b, pl->desc.b_scale, 1, s));
    if (!pl->desc.b_batched && gv.batch > 1 && x0_src) {
        // a sweep shares one start value: replicate it across the batch
        for (int bi = 0; bi < gv.batch; ++bi)
            NSOL_CHECK(nsol_scale_convert(ctx, gv.n, src_dtype, x0_src, gv.dtype, (char *)pl->x + (size_t)bi * gv.n * pl->esz,
                                          pl->desc.x0_scale, 1, s));
    } else {
        NSOL_CHECK(nsol_scale_convert(ctx, nv, src_dtype, x0_src, gv.dtype, pl->x, pl->desc.x0_scale, 1, s));
    }
    NSOL_CUDA(ctx, cudaMemcpyAsync(pl->xbar[0], pl->x, (size_t)nv * pl->esz, cudaMemcpyDeviceToDevice, s));
    for (int k = 0; k < gv.dim; ++k) NSOL_CUDA(ctx, cudaMemsetAsync(pl->p[0][k], 0, (size_t)nv * pl->esz, s));
    pl->cur = 0;
    pl->it = 0;
    pl->ready = true;
    pl->link_fresh = false;
    pl->chain_valid = false;
    return NSOL_OK;
}

extern "C" int nsol_pd_plan_reset_dev(nsol_pd_plan *pl, const void *b_dev, const void *x0_dev, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    if (!b_dev) return nsol_fail(pl->ctx, NSOL_EINVAL, "pd reset: b is NULL");
    NSOL_CHECK(nsol_bind_device(pl->ctx));
    return pd_reset_common(pl, pl->gv.dtype, b_dev, x0_dev ? x0_dev : b_dev, (cudaStream_t)s);
}

extern "C" int nsol_pd_plan_reset_host(nsol_pd_plan *pl, const double *b_host, const double *x0_host, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!b_host) return nsol_fail(ctx, NSOL_EINVAL, "pd reset: b is NULL");
    NSOL_CHECK(nsol_bind_device(ctx));
    const GridView &gv = pl->gv;
    cudaStream_t st = (cudaStream_t)s;
    const size_t nb = (size_t)gv.n * (pl->desc.b_batched ? gv.batch : 1);
    const bool same = (x0_host == nullptr) || (x0_host == b_host);
    const size_t nx0 = same ? 0 : (pl->desc.b_batched || gv.batch == 1 ? (size_t)gv.n * gv.batch : (size_t)gv.n);
    NSOL_CHECK(pd_ensure_stage(pl, (nb + nx0) * sizeof(double)));
    double *sb = (double *)pl->stage;
    double *sx = same ? sb : sb + nb;
    NSOL_CUDA(ctx, cudaMemcpyAsync(sb, b_host, nb * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!same) NSOL_CUDA(ctx, cudaMemcpyAsync(sx, x0_host, nx0 * sizeof(double), cudaMemcpyHostToDevice, st));
    return pd_reset_common(pl, NSOL_F64, sb, sx, st);
}

extern "C" int nsol_pd_plan_set_halo(nsol_pd_plan *pl, const void *xbar_above, const void *xbar_below, const void *p_below) {
    if (!pl) return NSOL_EINVAL;
    if ((xbar_below == nullptr) != (p_below == nullptr))
        return nsol_fail(pl->ctx, NSOL_EINVAL, "pd halo: xbar_below and p_below must be given together");
    if ((xbar_above || xbar_below) && pl->gv.dim < 2)
        return nsol_fail(pl->ctx, NSOL_EINVAL, "pd halo: slab decomposition needs dim >= 2");
    pl->halo_above = xbar_above;
    pl->halo_below = xbar_below;
    pl->halo_pz_below = p_below;
    return NSOL_OK;
}

static int pd_boundary_planes(nsol_pd_plan *pl, int which, const void **xbar_first, const void **xbar_last, const void **pz_last);

extern "C" int nsol_pd_plan_boundary_planes(nsol_pd_plan *pl, const void **xbar_first, const void **xbar_last, const void **pz_last) {
    return pd_boundary_planes(pl, pl ? pl->cur : 0, xbar_first, xbar_last, pz_last);
}
// the same planes of the buffers the running iteration writes (valid after nsol_pd_plan_iterate_part(plan, 1))
extern "C" int nsol_pd_plan_boundary_planes_next(nsol_pd_plan *pl, const void **xbar_first, const void **xbar_last, const void **pz_last) {
    return pd_boundary_planes(pl, pl ? (pl->cur ^ 1) : 0, xbar_first, xbar_last, pz_last);
}

static int pd_boundary_planes(nsol_pd_plan *pl, int which, const void **xbar_first, const void **xbar_last, const void **pz_last) {
    if (!pl) return NSOL_EINVAL;
    if (pl->gv.batch != 1) return nsol_fail(pl->ctx, NSOL_EINVAL, "pd boundary planes: batch must be 1");
    const GridView &gv = pl->gv;
    const size_t plane = (size_t)gv.nx * gv.ny * pl->esz;
    const char *xb = (const char *)pl->xbar[which];
    if (xbar_first) *xbar_first = xb;
    if (xbar_last) *xbar_last = xb + (size_t)(gv.nz - 1) * plane;
    if (pz_last) *pz_last = gv.comp_z >= 0 ? (const char *)pl->p[which][gv.comp_z] + (size_t)(gv.nz - 1) * plane : nullptr;
    return NSOL_OK;
}


// ---------------------------------------------------------------------------
// in-kernel halo exchange over peer memory
// ---------------------------------------------------------------------------
// Publishes the boundary planes of the CURRENT state (after a reset) as a new generation: waits
// until both neighbours have published the previous generation (so nobody still reads the slots
// about to be overwritten), copies the planes into the neighbours' receive slots and raises their
// flags -- the same protocol the iteration kernels follow for the states they produce.
struct LinkPublishArgs {
    const char *xbar_first, *xbar_last, *pz_last;        // boundary planes of the current state
    char *dst_below_xbar, *dst_above_xbar, *dst_above_pz;  // peer slots (NULL without that neighbour)
    const unsigned *flag_below, *flag_above;
    unsigned *peer_flag_below, *peer_flag_above;
    unsigned *count;                                     // local arrival counter (the "below" one is reused)
    int *error;
    unsigned long long plane_vec;                        // 16-byte words per plane
    unsigned long long timeout_ns;
    unsigned want, publish;
};

__global__ void __launch_bounds__(256) pd_link_publish_kernel(const LinkPublishArgs a) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned *)a.error + 2, 1u);      // diagnostics: publish kernels started
    if (a.flag_below) pd_link_wait(a.flag_below, a.want, a.error, a.timeout_ns);
    if (a.flag_above) pd_link_wait(a.flag_above, a.want, a.error, a.timeout_ns);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.plane_vec; i += stride) {
        if (a.dst_below_xbar) reinterpret_cast<float4 *>(a.dst_below_xbar)[i] = reinterpret_cast<const float4 *>(a.xbar_first)[i];
        if (a.dst_above_xbar) {
            reinterpret_cast<float4 *>(a.dst_above_xbar)[i] = reinterpret_cast<const float4 *>(a.xbar_last)[i];
            reinterpret_cast<float4 *>(a.dst_above_pz)[i] = reinterpret_cast<const float4 *>(a.pz_last)[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(a.count, 1u);
        if (done + 1u == gridDim.x) {
            atomicExch(a.count, 0u);
            __threadfence_system();
            if (a.peer_flag_below) st_release_sys(a.peer_flag_below, a.publish);
            if (a.peer_flag_above) st_release_sys(a.peer_flag_above, a.publish);
            atomicAdd((unsigned *)a.error + 3, 1u);                                        // diagnostics: ... finished
        }
    }
}

// everything queued on the plan's push stream happens before whatever is queued on s next
static int pd_push_join(nsol_pd_plan *pl, cudaStream_t s) {
    if (pl->push_stream && pl->push_count > 0)
        NSOL_CUDA(pl->ctx, cudaStreamWaitEvent(s, pl->push_ev_done[(pl->push_count - 1) & 7], 0));
    return NSOL_OK;
}

// The second stream of a linked plan and its events.  Called when the link is established -- while nothing spins on the device: the
// first launch into a fresh stream makes the driver set the stream up, and that has been seen to wait for the running kernel of the
// other stream (whose boundary CTAs were spinning for the neighbour's first publish: a time-out at generation 2 on 2 GPUs).  So the
// stream is created AND used once here.  The publish kernel also asks for the iteration kernel's shared-memory carve-out so that it
// can run beside it.
static int pd_push_setup(nsol_pd_plan *pl) {
    if (pl->push_stream) return NSOL_OK;
    nsol_ctx *ctx = pl->ctx;
    NSOL_CUDA(ctx, cudaFuncSetAttribute(pd_link_publish_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    NSOL_CUDA(ctx, cudaStreamCreateWithFlags(&pl->push_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) {
        NSOL_CUDA(ctx, cudaEventCreateWithFlags(&pl->push_ev_iter[i], cudaEventDisableTiming));
        NSOL_CUDA(ctx, cudaEventCreateWithFlags(&pl->push_ev_done[i], cudaEventDisableTiming));
    }
    LinkPublishArgs a = {};
    a.count = (unsigned *)(pl->link_block + LINK_COUNT_PUB);
    a.error = (int *)(pl->link_block + LINK_ERROR);
    pd_link_publish_kernel<<<32, 256, 0, pl->push_stream>>>(a);      // nothing to wait for, nothing to copy, no flag to raise
    for (int i = 0; i < 8; ++i) {
        NSOL_CUDA(ctx, cudaEventRecord(pl->push_ev_iter[i], pl->push_stream));
        NSOL_CUDA(ctx, cudaEventRecord(pl->push_ev_done[i], pl->push_stream));
    }
    NSOL_CUDA(ctx, cudaStreamSynchronize(pl->push_stream));
    return NSOL_OK;
}

// sides: 1 = towards the lower neighbour, 2 = towards the upper one, 3 = both (and the plan's generation counter advances unless
// advance is false); a one-sided call (pipelined solve: a boundary group has just been reset) publishes generation pl->link_pub
// without advancing it
static int pd_link_publish(nsol_pd_plan *pl, cudaStream_t s, int sides = 3, bool advance = true) {
    nsol_ctx *ctx = pl->ctx;
    const GridView &gv = pl->gv;
    const size_t plane = (size_t)gv.nx * gv.ny * pl->esz;
    const unsigned g = pl->link_pub;
    const bool below = pl->peer_below && (sides & 1), above = pl->peer_above && (sides & 2);
    if (!below && !above && sides != 3) return NSOL_OK;
    if (s != pl->push_stream) NSOL_CHECK(pd_push_join(pl, s));       // flags and slots are written in generation order
    const int wr = (int)(g & 1u);
    const char *xb = (const char *)pl->xbar[pl->cur];
    LinkPublishArgs a;
    a.xbar_first = xb;
    a.xbar_last = xb + (size_t)(gv.nz - 1) * plane;
    a.pz_last = (const char *)pl->p[pl->cur][gv.comp_z] + (size_t)(gv.nz - 1) * plane;
    a.dst_below_xbar = below ? pl->peer_below + link_slot(pl, wr, 0) : nullptr;
    a.dst_above_xbar = above ? pl->peer_above + link_slot(pl, wr, 1) : nullptr;
    a.dst_above_pz = above ? pl->peer_above + link_slot(pl, wr, 2) : nullptr;
    a.flag_below = below ? (const unsigned *)(pl->link_block + LINK_FLAG_BELOW) : nullptr;
    a.flag_above = above ? (const unsigned *)(pl->link_block + LINK_FLAG_ABOVE) : nullptr;
    a.peer_flag_below = below ? (unsigned *)(pl->peer_below + LINK_FLAG_ABOVE) : nullptr;
    a.peer_flag_above = above ? (unsigned *)(pl->peer_above + LINK_FLAG_BELOW) : nullptr;
    a.count = (unsigned *)(pl->link_block + LINK_COUNT_PUB);     // its own counter: a publish may run beside an iteration kernel
    a.error = (int *)(pl->link_block + LINK_ERROR);
    a.plane_vec = plane / 16;
    a.timeout_ns = (unsigned long long)(ctx->link_timeout_ms > 0 ? ctx->link_timeout_ms : NSOL_LINK_TIMEOUT_MS) * 1000000ull;
    a.want = g;               // neighbours have published generation g - 1 (trivially true for g = 0)
    a.publish = g + 1u;
    pd_link_publish_kernel<<<32, 256, 0, s>>>(a);
    NSOL_LAUNCH_CHECK(ctx);
    if (sides == 3 && advance) {
        pl->link_pub = g + 1u;
        pl->link_fresh = true;
    }
    return NSOL_OK;
}

extern "C" int nsol_pd_plan_link_create(nsol_pd_plan *pl, void **block_dev, size_t *block_bytes) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    const GridView &gv = pl->gv;
    if (gv.dim < 2) return nsol_fail(ctx, NSOL_EINVAL, "pd link: slab decomposition needs dim >= 2");
    if (gv.batch != 1) return nsol_fail(ctx, NSOL_EINVAL, "pd link: batch must be 1");
    const size_t plane = (size_t)gv.nx * gv.ny * pl->esz;
    if (plane % 16) return nsol_fail(ctx, NSOL_EINVAL, "pd link: a plane must be a multiple of 16 bytes (nx*ny = %lld)", (long long)gv.nx * gv.ny);
    NSOL_CHECK(nsol_bind_device(ctx));
    if (!pl->link_block) {
        pl->link_plane = (plane + 255) / 256 * 256;
        const size_t bytes = LINK_HEADER + 6 * pl->link_plane;
        NSOL_CUDA(ctx, cudaMalloc((void **)&pl->link_block, bytes));
        NSOL_CUDA(ctx, cudaMemset(pl->link_block, 0, bytes));
        pl->bytes += bytes;
    }
    if (block_dev) *block_dev = pl->link_block;
    if (block_bytes) *block_bytes = LINK_HEADER + 6 * pl->link_plane;
    return NSOL_OK;
}

extern "C" int nsol_pd_plan_link_ipc_handle(nsol_pd_plan *pl, void *handle64) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!handle64) return nsol_fail(ctx, NSOL_EINVAL, "pd link: handle buffer is NULL");
    NSOL_CHECK(nsol_pd_plan_link_create(pl, nullptr, nullptr));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    NSOL_CUDA(ctx, cudaIpcGetMemHandle(&h, pl->link_block));
    memcpy(handle64, &h, 64);
    return NSOL_OK;
}

static int pd_link_set_peers(nsol_pd_plan *pl, char *below, bool below_ipc, char *above, bool above_ipc) {
    nsol_ctx *ctx = pl->ctx;
    if (!pl->link_block) return nsol_fail(ctx, NSOL_ESTATE, "pd link: call nsol_pd_plan_link_create first");
    if (pl->link_on) return nsol_fail(ctx, NSOL_ESTATE, "pd link: already connected");
    pl->peer_below = below;
    pl->peer_below_ipc = below_ipc;
    pl->peer_above = above;
    pl->peer_above_ipc = above_ipc;
    pl->link_on = (below != nullptr) || (above != nullptr);
    pl->link_fresh = false;
    pl->link_pub = 0;
    if (pl->link_on && ctx->pd_push == 1) NSOL_CHECK(pd_push_setup(pl));
    return NSOL_OK;
}

extern "C" int nsol_pd_plan_link_connect(nsol_pd_plan *pl, void *block_below, void *block_above) {
    if (!pl) return NSOL_EINVAL;
    return pd_link_set_peers(pl, (char *)block_below, false, (char *)block_above, false);
}

extern "C" int nsol_pd_plan_link_open(nsol_pd_plan *pl, const void *handle_below64, const void *handle_above64) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!pl->link_block) return nsol_fail(ctx, NSOL_ESTATE, "pd link: call nsol_pd_plan_link_create first");
    if (pl->link_on) return nsol_fail(ctx, NSOL_ESTATE, "pd link: already connected");
    NSOL_CHECK(nsol_bind_device(ctx));
    void *below = nullptr, *above = nullptr;
    cudaIpcMemHandle_t h;
    if (handle_below64) {
        memcpy(&h, handle_below64, 64);
        NSOL_CUDA(ctx, cudaIpcOpenMemHandle(&below, h, cudaIpcMemLazyEnablePeerAccess));
    }
    if (handle_above64) {
        memcpy(&h, handle_above64, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&above, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            if (below) cudaIpcCloseMemHandle(below);
            return nsol_fail(ctx, NSOL_ECUDA, "pd link: cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e));
        }
    }
    return pd_link_set_peers(pl, (char *)below, below != nullptr, (char *)above, above != nullptr);
}

// synchronises the stream and reports a timed-out wait of any launch so far
extern "C" int nsol_pd_plan_link_status(nsol_pd_plan *pl, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!pl->link_block) return NSOL_OK;
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CHECK(pd_push_join(pl, (cudaStream_t)s));
    int err = 0;
    NSOL_CUDA(ctx, cudaMemcpyAsync(&err, pl->link_block + LINK_ERROR, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)s));
    NSOL_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)s));
    if (err) {
        unsigned hdr[64] = {0};
        cudaMemcpy(hdr, pl->link_block, sizeof(hdr), cudaMemcpyDeviceToHost);
        return nsol_fail(ctx, NSOL_ENCCL, "pd link: a wait for a neighbour's halo timed out (neighbour not iterating in lockstep?) "
                         "[first missing generation %u (own publish kernels started %u, finished %u at that moment; now %u, %u); now: flag_below %u flag_above %u count_below %u count_above %u count_pub %u; generation %u, %llu publish kernels]",
                         hdr[LINK_ERROR / 4 + 1], hdr[LINK_ERROR / 4 + 4], hdr[LINK_ERROR / 4 + 5], hdr[LINK_ERROR / 4 + 2], hdr[LINK_ERROR / 4 + 3], hdr[LINK_FLAG_BELOW / 4], hdr[LINK_FLAG_ABOVE / 4], hdr[LINK_COUNT_BELOW / 4], hdr[LINK_COUNT_ABOVE / 4],
                         hdr[LINK_COUNT_PUB / 4], pl->link_pub, pl->push_count);
    }
    return NSOL_OK;
}

// Query mode: the launch helpers below only make sure the kernel they would launch is LOADED (CUDA loads kernels lazily, and loading
// one may wait for every running kernel of the context -- which must not happen behind a kernel that spins on a neighbour's flag:
// the pipelined solve through linked z-slabs preloads what it is going to launch, pd_preload).
static thread_local bool g_pd_query_only = false;
// Programmatic dependent launch ("pd_pdl" knob: 0 on, 2 off): the iteration kernels call griddepcontrol.launch_dependents when they
// start and griddepcontrol.wait before their first read of solver state, so the CTAs of iteration k + 1 are launched (shared memory
// carved, mbarriers initialised) while the last wave of iteration k drains, and start the moment it has completed -- the launch gap
// and ramp of a 0.24 ms kernel on a 64-plane z-slab are a few per cent of it.
static thread_local bool g_pd_pdl = true;
#define NSOL_PD_LAUNCH(KERNEL, GRID, BLOCK, SMEM, STREAM, ARGS)                                                   \
    do {                                                                                                          \
        if (g_pd_query_only) {                                                                                    \
            cudaFuncAttributes fa__;                                                                              \
            cudaFuncGetAttributes(&fa__, KERNEL);                                                                 \
        } else if (g_pd_pdl) {                                                                                    \
            cudaLaunchConfig_t cfg__ = {};                                                                        \
            cfg__.gridDim = GRID;                                                                                 \
            cfg__.blockDim = BLOCK;                                                                               \
            cfg__.dynamicSmemBytes = SMEM;                                                                        \
            cfg__.stream = STREAM;                                                                                \
            cudaLaunchAttribute at__[1];                                                                          \
            at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                      \
            at__[0].val.programmaticStreamSerializationAllowed = 1;                                               \
            cfg__.attrs = at__;                                                                                   \
            cfg__.numAttrs = 1;                                                                                   \
            cudaLaunchKernelEx(&cfg__, KERNEL, ARGS);                                                             \
        } else {                                                                                                  \
            KERNEL<<<GRID, BLOCK, SMEM, STREAM>>>(ARGS);                                                          \
        }                                                                                                         \
    } while (0)

template <typename T, int VEC, bool HAS_Y>
static void pd_launch_rd(const nsol_pd_plan *pl, const PdArgs<T> &a, dim3 grid, dim3 block, size_t smem, cudaStream_t s) {
    const int reg = pl->desc.reg, data = pl->desc.data;
    const bool link = pl->link_on;
    // unit spacing (every BASELINE configuration): the w * a products are exact and dropped (vector kernels only)
    constexpr bool CAN_UNIT = VEC > 1;
    const bool unit = CAN_UNIT && a.wx == T(1) && (!HAS_Y || a.wy == T(1)) && (!a.has_z || a.wz == T(1));
#define NSOL_PD_CASE(R, D)                                                                                   \
    if (reg == R && data == D) {                                                                             \
        if (unit) {                                                                                          \
            if (link) NSOL_PD_LAUNCH((pd_iter_kernel<T, VEC, HAS_Y, R, D, true, CAN_UNIT>), grid, block, smem, s, a);   \
            else NSOL_PD_LAUNCH((pd_iter_kernel<T, VEC, HAS_Y, R, D, false, CAN_UNIT>), grid, block, smem, s, a);       \
        } else {                                                                                             \
            if (link) NSOL_PD_LAUNCH((pd_iter_kernel<T, VEC, HAS_Y, R, D, true, false>), grid, block, smem, s, a);      \
            else NSOL_PD_LAUNCH((pd_iter_kernel<T, VEC, HAS_Y, R, D, false, false>), grid, block, smem, s, a);          \
        }                                                                                                    \
        return;                                                                                              \
    }
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L1)
#undef NSOL_PD_CASE
}

// bulk-async (TMA) staged variant: 3-D, vector path only
template <typename T, int VEC, int R, int D, bool LINK, bool UNIT>
static int pd_launch_bulk_one(const nsol_pd_plan *pl, const PdArgs<T> &a, dim3 grid, dim3 block, size_t smem, cudaStream_t s) {
    // the opt-in is per device and per kernel instantiation; remember the largest size set on each device
    static size_t configured[64] = {0};
    const int dev = pl->ctx->device & 63;
    if (smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pd_iter_bulk_kernel<T, VEC, R, D, LINK, UNIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return nsol_fail(pl->ctx, NSOL_ECUDA, "pd bulk: smem opt-in %zu -> %s", smem, cudaGetErrorString(e));
        configured[dev] = smem;
    }
    NSOL_PD_LAUNCH((pd_iter_bulk_kernel<T, VEC, R, D, LINK, UNIT>), grid, block, smem, s, a);
    return NSOL_OK;
}

template <typename T, int VEC>
static int pd_launch_bulk(const nsol_pd_plan *pl, const PdArgs<T> &a, dim3 grid, dim3 block, size_t smem, cudaStream_t s) {
    const int reg = pl->desc.reg, data = pl->desc.data;
    const bool link = pl->link_on;
    const bool unit = a.wx == T(1) && a.wy == T(1) && a.wz == T(1);
#define NSOL_PD_CASE(R, D)                                                                              \
    if (reg == R && data == D) {                                                                        \
        if (link) return unit ? pd_launch_bulk_one<T, VEC, R, D, true, true>(pl, a, grid, block, smem, s)    \
                              : pd_launch_bulk_one<T, VEC, R, D, true, false>(pl, a, grid, block, smem, s);  \
        return unit ? pd_launch_bulk_one<T, VEC, R, D, false, true>(pl, a, grid, block, smem, s)         \
                    : pd_launch_bulk_one<T, VEC, R, D, false, false>(pl, a, grid, block, smem, s);       \
    }
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L1)
#undef NSOL_PD_CASE
    return NSOL_EINVAL;
}

// persistent cooperative launch of n iterations (2-D / 1-D vector path, no z-slab neighbours)
template <typename T, int VEC, int R, int D, bool UNIT>
static int pd_launch_persist_one(nsol_pd_plan *pl, PdArgs<T> a, dim3 grid, dim3 block, int n, cudaStream_t s) {
    nsol_ctx *ctx = pl->ctx;
    static int per_sm[64] = {0};      // co-resident CTAs per SM of this instantiation, per device
    const int dev = ctx->device & 63;
    if (per_sm[dev] == 0) {
        int v = 0;
        NSOL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, pd_iter_persist_kernel<T, VEC, R, D, UNIT>, (int)block.x, 0));
        per_sm[dev] = v > 0 ? v : -1;
    }
    if (per_sm[dev] < 0) return NSOL_ESTATE;
    unsigned vgx = grid.x, vgz = grid.z;
    const long long total = (long long)vgx * vgz;
    // as many CTAs as are co-resident: measured on B200 (profiles/r2_latency_configs.md) the iteration time falls with the CTA
    // count until every virtual block has its own CTA (256^2: flat from 74 CTAs at 4.9 us; 1024^2: 27 us with 148 CTAs, 11.7 us
    // with 592) -- the phases are bound by the latency of their dependent L2 accesses, not by the cost of grid.sync()
    long long cap = (long long)per_sm[dev] * ctx->sm_count;
    if (ctx->pd_persist_blocks > 0 && ctx->pd_persist_blocks < cap) cap = ctx->pd_persist_blocks;
    const unsigned blocks = (unsigned)(total < cap ? total : cap);
    void *params[] = {(void *)&a, (void *)&n, (void *)&vgx, (void *)&vgz};
    NSOL_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)pd_iter_persist_kernel<T, VEC, R, D, UNIT>, dim3(blocks), block, params, 0, s));
    ctx->launches++;
    return NSOL_OK;
}

template <typename T, int VEC>
static int pd_launch_persist(nsol_pd_plan *pl, const PdArgs<T> &a, dim3 grid, dim3 block, int n, cudaStream_t s) {
    const int reg = pl->desc.reg, data = pl->desc.data;
    const bool unit = a.wx == T(1) && (!a.has_z || a.wz == T(1));
#define NSOL_PD_CASE(R, D)                                                                           \
    if (reg == R && data == D)                                                                       \
        return unit ? pd_launch_persist_one<T, VEC, R, D, true>(pl, a, grid, block, n, s)            \
                    : pd_launch_persist_one<T, VEC, R, D, false>(pl, a, grid, block, n, s);
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TV, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_HUBER, NSOL_DATA_L1)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L2)
    NSOL_PD_CASE(NSOL_REG_TK1, NSOL_DATA_L1)
#undef NSOL_PD_CASE
    return NSOL_EINVAL;
}

// rows per CTA (3-D only) and planes per z-chunk
static void pd_tiling(const nsol_ctx *ctx, const GridView &gv, int vecw, int *ty_out, int *zc_out, bool link = false) {
    const bool has_y = gv.comp_y >= 0;
    int ty = 1;
    if (has_y) {
        // thin linked z-slabs (strong scaling: 64 planes per GPU at N = 8): 4 rows per CTA -- half as many boundary CTAs paying the
        // system-scope fence; measured on 2 x 64 planes 0.2454 vs 0.2494 ms per launch (profiles/r2_thin_slabs.md)
        ty = ctx->pd_ty ? ctx->pd_ty : ((link && gv.nz <= 96) ? 4 : 2);
        if (ty < 1) ty = 1;
        if (ty > 8) ty = 8;
    }
    const int tile_w = (has_y ? 32 : 128) * vecw;
    int zc = ctx->pd_zc;
    if (zc <= 0) {
        // measured on B200 at 512^3: 16 planes per chunk for the register-pipelined kernels (profiles/r1_tuning.md),
        // 8 for the TMA bulk-async kernel since its instruction count was cut (profiles/r1_tuning_v2.log: 4 planes
        // 0.950, 6 0.986, 8 0.991, 10 0.985, 12 0.978, 16 0.969 of the roofline, 24 0.958, 32 0.938, 128 0.890);
        // shorter chunks only when the volume would otherwise give fewer than ~4 CTAs per SM
        const bool bulk = has_y && vecw > 1 && (ctx->pd_variant == 2 || (ctx->pd_variant == 0 && gv.dtype == NSOL_F64 && NSOL_PD_DEFAULT_VARIANT_F64 == 2) ||
                                                (ctx->pd_variant == 0 && gv.dtype == NSOL_F32 && NSOL_PD_DEFAULT_VARIANT_F32 == 2));
        zc = bulk ? 8 : 16;
        const long long tiles = (long long)((gv.nx + tile_w - 1) / tile_w) * (has_y ? (gv.ny + ty - 1) / ty : 1) * gv.batch;
        while (zc > 4 && tiles * ((gv.nz + zc - 1) / zc) < (long long)ctx->sm_count * 4) zc /= 2;
        // thin z-slabs with the in-kernel halo exchange: keep the two boundary chunks (whose CTAs end
        // with a system-scope fence) a small share of the slab; measured on a 64-plane slab: 8 planes
        if (link && zc > 8 && gv.nz < 8 * zc) zc = 8;
    }
    if (zc > gv.nz) zc = gv.nz;
    if (zc < 1) zc = 1;
    *ty_out = ty;
    *zc_out = zc;
}

// part: 0 = whole iteration; 1 = only the first and last z-chunk (the planes a z-slab neighbour
// needs), no state advance; 2 = the remaining interior chunks, then advance.
// rg (pipelined solve, nsol_pd_plan_solve_host): iteration rg->it (0-based) on the z-chunks [rg->c0, rg->c1) only, reading the
// buffers of parity it & 1; the plan's own counters do not move.
struct PdRange {
    int c0, c1, it;
    int zc;        // > 0: planes per z-chunk of this launch (the result does not depend on the chunking)
};

template <typename T, int VECW>
static int pd_launch_iteration(nsol_pd_plan *pl, cudaStream_t s, int part = 0, int persist_n = 0, const PdRange *rg = nullptr) {
    nsol_ctx *ctx = pl->ctx;
    const GridView &gv = pl->gv;
    if (rg && (part != 0 || persist_n > 1)) return nsol_fail(ctx, NSOL_ESTATE, "pd: chunk-range launch in split / persistent mode");
    const int cur = rg ? (rg->it & 1) : pl->cur, nxt = cur ^ 1;
    PdArgs<T> a;
    a.xbar_in = (const T *)pl->xbar[cur];
    a.xbar_out = (T *)pl->xbar[nxt];
    a.x = (T *)pl->x;
    a.b = (const T *)pl->b;
    a.px_in = (const T *)pl->p[cur][gv.comp_x];
    a.px_out = (T *)pl->p[nxt][gv.comp_x];
    a.py_in = gv.comp_y >= 0 ? (const T *)pl->p[cur][gv.comp_y] : nullptr;
    a.py_out = gv.comp_y >= 0 ? (T *)pl->p[nxt][gv.comp_y] : nullptr;
    a.pz_in = gv.comp_z >= 0 ? (const T *)pl->p[cur][gv.comp_z] : nullptr;
    a.pz_out = gv.comp_z >= 0 ? (T *)pl->p[nxt][gv.comp_z] : nullptr;
    a.sched = pl->sched;
    a.halo_xbar_above = (const T *)pl->halo_above;
    a.halo_xbar_below = (const T *)pl->halo_below;
    a.halo_pz_below = (const T *)pl->halo_pz_below;
    a.flag_below = a.flag_above = nullptr;
    a.peer_flag_below = a.peer_flag_above = nullptr;
    a.count_below = a.count_above = nullptr;
    a.push_below_xbar = a.push_above_xbar = a.push_above_pz = nullptr;
    a.link_error = nullptr;
    a.want = a.publish = 0;
    a.front_chunks = 0;
    a.chain_done = nullptr;
    a.chain_need = 0;
    a.chain = 0;
    a.link_timeout_ns = (unsigned long long)(ctx->link_timeout_ms > 0 ? ctx->link_timeout_ms : NSOL_LINK_TIMEOUT_MS) * 1000000ull;
    if (pl->link_on) {
        if (part != 0) return nsol_fail(ctx, NSOL_ESTATE, "pd: the split iteration is not available with the in-kernel halo exchange");
        // this iteration consumes generation g = link_pub - 1 and publishes generation link_pub; a chunk-range launch of the
        // pipelined solve counts from the generation of the solve's start state (every boundary has its own pace there)
        const unsigned link_pub = rg ? pl->pipe_g0 + 1u + (unsigned)rg->it : pl->link_pub;
        const unsigned g = link_pub - 1;
        const int rd = (int)(g & 1u), wr = (int)(link_pub & 1u);
        a.link_error = (int *)(pl->link_block + LINK_ERROR);
        a.want = g + 1u;
        a.publish = link_pub + 1u;
        if (pl->peer_below) {
            a.halo_xbar_below = (const T *)(pl->link_block + link_slot(pl, rd, 1));
            a.halo_pz_below = (const T *)(pl->link_block + link_slot(pl, rd, 2));
            a.flag_below = (const unsigned *)(pl->link_block + LINK_FLAG_BELOW);
            a.count_below = (unsigned *)(pl->link_block + LINK_COUNT_BELOW);
            a.push_below_xbar = (T *)(pl->peer_below + link_slot(pl, wr, 0));      // their xbar_above
            a.peer_flag_below = (unsigned *)(pl->peer_below + LINK_FLAG_ABOVE);    // I am their upper neighbour
        }
        if (pl->peer_above) {
            a.halo_xbar_above = (const T *)(pl->link_block + link_slot(pl, rd, 0));
            a.flag_above = (const unsigned *)(pl->link_block + LINK_FLAG_ABOVE);
            a.count_above = (unsigned *)(pl->link_block + LINK_COUNT_ABOVE);
            a.push_above_xbar = (T *)(pl->peer_above + link_slot(pl, wr, 1));      // their xbar_below
            a.push_above_pz = (T *)(pl->peer_above + link_slot(pl, wr, 2));        // their pz_below
            a.peer_flag_above = (unsigned *)(pl->peer_above + LINK_FLAG_BELOW);    // I am their lower neighbour
        }
    }
    a.n = gv.n;
    a.b_stride = pl->desc.b_batched ? gv.n : 0;
    a.nx = gv.nx;
    a.ny = gv.ny;
    a.nz = gv.nz;
    a.has_z = gv.comp_z >= 0;
    a.wx = (T)gv.w[gv.comp_x];
    a.wy = gv.comp_y >= 0 ? (T)gv.w[gv.comp_y] : T(0);
    a.wz = gv.comp_z >= 0 ? (T)gv.w[gv.comp_z] : T(0);
    a.it = rg ? rg->it : pl->it;
    a.batch = gv.batch;

    const bool has_y = gv.comp_y >= 0;
    dim3 block, grid;
    size_t smem = 0;
    int ty = 1, zc = 1;
    pd_tiling(ctx, gv, VECW, &ty, &zc, pl->link_on);
    if (rg && rg->zc > 0) zc = std::min(rg->zc, gv.nz);
    if (has_y) {
        block = dim3(32, ty, 1);
        smem = (size_t)(2 * (ty + 2) + 2 * (ty + 1)) * 32 * VECW * sizeof(T);
    } else {
        block = dim3(128, 1, 1);
    }
    const int tile_w = block.x * VECW;
    grid.x = (gv.nx + tile_w - 1) / tile_w;
    grid.y = has_y ? (gv.ny + ty - 1) / ty : 1;
    a.zc = zc;
    a.nchunks = (gv.nz + zc - 1) / zc;
    a.nsel = a.nchunks;
    a.chunk_first = 0;
    a.chunk_stride = 1;
    if (part != 0) {
        if (a.nchunks < 3) return nsol_fail(ctx, NSOL_ESTATE, "pd: split iteration needs at least 3 z-chunks (nz=%d, zc=%d)", gv.nz, zc);
        if (part == 1) {
            a.nsel = 2;
            a.chunk_stride = a.nchunks - 1;
        } else {
            a.nsel = a.nchunks - 2;
            a.chunk_first = 1;
        }
    }
    if (rg) {
        if (rg->c0 < 0 || rg->c1 > a.nchunks || rg->c0 >= rg->c1) return nsol_fail(ctx, NSOL_EINVAL, "pd: bad chunk range [%d, %d) of %d", rg->c0, rg->c1, a.nchunks);
        a.chunk_first = rg->c0;
        a.nsel = rg->c1 - rg->c0;
        // a boundary that is not part of the range neither waits nor publishes in this launch
        if (rg->c0 > 0) {
            a.flag_below = nullptr;
            a.count_below = nullptr;
            a.push_below_xbar = nullptr;
            a.peer_flag_below = nullptr;
        }
        if (rg->c1 < a.nchunks) {
            a.flag_above = nullptr;
            a.count_above = nullptr;
            a.push_above_xbar = a.push_above_pz = nullptr;
            a.peer_flag_above = nullptr;
        }
    }
    a.front_chunks = (pl->link_on && a.nchunks >= 3 && !rg) ? 1 : 0;
    // Link mode, whole-slab iteration: the boundary CTAs do not push their planes to the neighbours (peer stores + a system-scope
    // fence per CTA, while the CTA holds its SM slot -- a quarter of all CTAs on a 64-plane slab); a publish kernel on the plan's
    // second stream copies the three boundary planes and raises the flags once the iteration kernel has completed.  The boundary
    // chunks are scheduled LAST, so the neighbours' planes of the previous iteration -- published ~20 us after that iteration
    // ended -- have long arrived when they are needed ("pd_push" knob: 2 = push from inside the kernel as before).
    // EXPERIMENT, off unless "pd_push" = 1: on 2 real GPUs it still runs into halo-wait time-outs (first a race on a re-recorded
    // event -- fixed by the event rings --, then one at the first solve after a device-wide synchronisation; not understood), so the
    // boundary CTAs keep pushing from inside the kernel by default.  Single-GPU emulation and the small 2-GPU API check pass.
    const bool ext_push = pl->link_on && !rg && part == 0 && ctx->pd_push == 1 && !g_pd_query_only;
    if (ext_push) {
        a.push_below_xbar = a.push_above_xbar = a.push_above_pz = nullptr;
        a.peer_flag_below = a.peer_flag_above = nullptr;
        a.count_below = a.count_above = nullptr;
        a.front_chunks = a.nchunks >= 3 ? 2 : 0;
        NSOL_CHECK(pd_push_setup(pl));
        // this launch overwrites the buffers the publish kernel of two iterations ago read
        if (pl->push_count >= 2) NSOL_CUDA(ctx, cudaStreamWaitEvent(s, pl->push_ev_done[(pl->push_count - 2) & 7], 0));
    } else if (pl->link_on && !g_pd_query_only) {
        NSOL_CHECK(pd_push_join(pl, s));        // in-kernel pushes of this launch come after every queued publish kernel
    }
    const long long gz = (long long)a.nsel * gv.batch;
    if (gz > 65535) return nsol_fail(ctx, NSOL_EINVAL, "pd: nchunks*batch = %lld exceeds the grid limit; raise pd_zc", gz);
    grid.z = (unsigned)gz;
    if (persist_n > 1) {
        // small 2-D / 1-D problem (the caller checked): all persist_n iterations in one cooperative launch
        if constexpr (VECW > 1) {
            int rc = pd_launch_persist<T, VECW>(pl, a, grid, block, persist_n, s);
            if (rc == NSOL_OK) {
                pl->cur = (persist_n & 1) ? nxt : cur;
                pl->it += persist_n;
            }
            return rc;
        }
        return NSOL_ESTATE;
    }
    // No programmatic dependent launch together with the publish kernel: CTAs of iteration k + 1 launched early sit on the SMs
    // (blocked in griddepcontrol.wait) while the boundary CTAs of iteration k spin for the neighbour's planes -- and the publish
    // kernel of iteration k - 1 on the second stream, which the neighbour is waiting for in turn, finds no free SM (observed as a
    // halo-wait time-out on 2 GPUs).
    g_pd_pdl = ctx->pd_pdl != 2 && !ext_push;
    // iteration chaining (pd_chain_begin): whole-volume launches of one 3-D volume, back to back.  OFF unless "pd_chain" = 1: measured
    // on B200 it LOSES (512^3 float64: 2.138 vs 1.895 ms per launch, float32 96.6 vs 92.9 ms per step; profiles/r2_thin_slabs.md) --
    // the per-CTA release fence waits for the CTA's stores to be acknowledged by a saturated HBM write queue while the CTA holds its
    // SM slot, which costs more than the overlapped head / tail of the launches wins.  Kept as a measured, parity-tested experiment.
    const bool chain_ok = !rg && part == 0 && has_y && gv.batch == 1 && g_pd_pdl && ctx->pd_chain == 1 && !g_pd_query_only;
    if (chain_ok) {
        if (pl->chain_chunks != a.nchunks || pl->chain_zc != zc) {
            if (pl->chain_done) {
                NSOL_CUDA(ctx, cudaStreamSynchronize(s));
                NSOL_CUDA(ctx, cudaFree(pl->chain_done));
                pl->chain_done = nullptr;
            }
            NSOL_CUDA(ctx, cudaMalloc((void **)&pl->chain_done, sizeof(unsigned long long) * (size_t)a.nchunks));
            pl->chain_chunks = a.nchunks;
            pl->chain_zc = zc;
            pl->chain_valid = false;
        }
        if (!pl->chain_valid) {
            NSOL_CUDA(ctx, cudaMemsetAsync(pl->chain_done, 0, sizeof(unsigned long long) * (size_t)a.nchunks, s));
            pl->chain_gen = 0;
            pl->chain_valid = true;
        }
        a.chain_done = pl->chain_done;
        a.chain = pl->chain_gen > 0 ? 1 : 0;        // the first launch of a chain waits for whatever came before it
        a.chain_need = pl->chain_gen * (unsigned long long)(grid.x * grid.y);
    } else if (!g_pd_query_only) {
        pl->chain_valid = false;
    }
    // kernel variant: 1 = register-pipelined loads (LDG), 2 = TMA bulk-async staged tiles;
    // default: bulk for float64 3-D volumes (issue-bound with LDG), LDG otherwise
    int variant = ctx->pd_variant;
    if (variant == 0) variant = (sizeof(T) == 8) ? NSOL_PD_DEFAULT_VARIANT_F64 : NSOL_PD_DEFAULT_VARIANT_F32;
    if (variant == 2 && has_y && VECW > 1) {
        smem = PdBulkLayout<T, (VECW > 1 ? VECW : 2)>::bytes(ty, NSOL_PD_STAGES);
        NSOL_CHECK((pd_launch_bulk<T, (VECW > 1 ? VECW : 2)>(pl, a, grid, block, smem, s)));
    } else if (has_y) {
        pd_launch_rd<T, VECW, true>(pl, a, grid, block, smem, s);
    } else {
        pd_launch_rd<T, VECW, false>(pl, a, grid, block, smem, s);
    }
    if (g_pd_query_only) return NSOL_OK;
    NSOL_LAUNCH_CHECK(ctx);
    if (chain_ok) pl->chain_gen += 1;
    if (part != 1 && !rg) {
        pl->cur = nxt;
        pl->it += 1;
        if (ext_push) {
            // publish the state this launch produces (generation link_pub, slot link_pub & 1 -- what the boundary CTAs would have done)
            NSOL_CUDA(ctx, cudaEventRecord(pl->push_ev_iter[pl->push_count & 7], s));
            NSOL_CUDA(ctx, cudaStreamWaitEvent(pl->push_stream, pl->push_ev_iter[pl->push_count & 7], 0));
            NSOL_CHECK(pd_link_publish(pl, pl->push_stream, 3, false));
            NSOL_CUDA(ctx, cudaEventRecord(pl->push_ev_done[pl->push_count & 7], pl->push_stream));
            pl->push_count += 1;
        }
        if (pl->link_on) pl->link_pub += 1;
    }
    return NSOL_OK;
}

// n iterations of a 2-D problem with the temporal-blocking kernel: ceil(n / K) launches.  NSOL_ESTATE: not applicable here.
template <typename T>
static int pd_iterate_tb2d(nsol_pd_plan *pl, int n, cudaStream_t st) {
    nsol_ctx *ctx = pl->ctx;
    const GridView &gv = pl->gv;
    // region: 64 columns x 32 rows (512 threads, two rows per thread); "pd_tb_nr" = 1: 64 x 16, 4: 64 x 32 with 256 threads
    const long long pixels = (long long)gv.nx * gv.nz * gv.batch;
    const bool small = pixels < (long long)ctx->sm_count * 4 * 56 * 24;
    const int nr = ctx->pd_tb_nr > 0 ? ctx->pd_tb_nr : 2;
    const int rows = pd_tb2d_rows(nr);
    if (rows == 0) return nsol_fail(ctx, NSOL_EINVAL, "pd_tb_nr must be 1, 2 or 4");
    // iterations per pass, measured on B200 (profiles/r2_tb2d.md): 4 where the kernel is bound by its arithmetic (the recomputed
    // rings grow with K), 8 for a small single image, which is bound by launches and dependent L2 round trips
    int K = ctx->pd_tb_k > 0 ? ctx->pd_tb_k : (small ? 8 : 4);
    K = std::min(K, (rows - 2) / 2);
    const int tw = 64 - 2 * K, th = rows - 2 * K;
    const long long tiles_z = (gv.nz + th - 1) / th;
    if (tiles_z > 65535 || gv.batch > 65535) return NSOL_ESTATE;
    if (!pl->x_alt) {
        cudaError_t e = nsol_plan_alloc(ctx, &pl->x_alt, (size_t)gv.n * gv.batch * pl->esz);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return NSOL_ESTATE;      // no room for the second x array: one pass per iteration
        }
        pl->bytes += (size_t)gv.n * gv.batch * pl->esz;
    }
    PdTbArgs<T> a;
    a.b = (const T *)pl->b;
    a.sched = pl->sched;
    a.n = gv.n;
    a.b_stride = pl->desc.b_batched ? gv.n : 0;
    a.nx = gv.nx;
    a.nz = gv.nz;
    a.batch = gv.batch;
    a.halo = K;
    a.tw = tw;
    a.th = th;
    a.wx = (T)gv.w[gv.comp_x];
    a.wz = (T)gv.w[gv.comp_z];
    const dim3 grid((gv.nx + tw - 1) / tw, (unsigned)tiles_z, gv.batch);
    for (int done = 0; done < n;) {
        const int k = std::min(K, n - done);
        const int cur = pl->cur, nxt = cur ^ 1;
        a.xbar_in = (const T *)pl->xbar[cur];
        a.xbar_out = (T *)pl->xbar[nxt];
        a.px_in = (const T *)pl->p[cur][gv.comp_x];
        a.px_out = (T *)pl->p[nxt][gv.comp_x];
        a.pz_in = (const T *)pl->p[cur][gv.comp_z];
        a.pz_out = (T *)pl->p[nxt][gv.comp_z];
        a.x_in = (const T *)pl->x;
        a.x_out = (T *)pl->x_alt;
        a.it = pl->it;
        a.ksub = k;
        pl->chain_valid = false;
        NSOL_CHECK(pd_tb2d_launch<T>(ctx, pl->desc.reg, pl->desc.data, nr, a, grid, st));
        NSOL_LAUNCH_CHECK(ctx);
        std::swap(pl->x, pl->x_alt);
        pl->cur = nxt;
        pl->it += k;
        done += k;
    }
    return NSOL_OK;
}

extern "C" int nsol_pd_plan_iterate(nsol_pd_plan *pl, int n, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!pl->ready) return nsol_fail(ctx, NSOL_ESTATE, "pd iterate: plan has not been reset with an observation");
    if (n < 0) return nsol_fail(ctx, NSOL_EINVAL, "pd iterate: n must be >= 0");
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CHECK(pd_ensure_schedule(pl, pl->it + n));
    const GridView &gv = pl->gv;
    cudaStream_t st = (cudaStream_t)s;
    const int vecw = gv.dtype == NSOL_F32 ? 4 : 2;
    const bool vec_ok = (gv.nx % vecw) == 0;
    if (pl->link_on && !pl->link_fresh) NSOL_CHECK(pd_link_publish(pl, st));   // also for n == 0 (publish only)
    // 2-D problems: K iterations per pass over the state in shared-memory tiles (csrc/pd_tb2d.cuh; "pd_tb" tuning knob: 0 auto,
    // 1 whenever possible, 2 never).  Single images (BASELINE configs 1, 2) lose 1 - 1/K of their dependent L2 round trips and
    // launches, batched sweeps (config 5) 1 - 1/K of their HBM traffic.
    if (gv.dim == 2 && gv.comp_z >= 0 && gv.comp_y < 0 && !pl->link_on && !pl->halo_above && !pl->halo_below && ctx->pd_tb != 2 &&
        (n >= 2 || ctx->pd_tb == 1) && n >= 1) {
        int rc = gv.dtype == NSOL_F32 ? pd_iterate_tb2d<float>(pl, n, st) : pd_iterate_tb2d<double>(pl, n, st);
        if (rc != NSOL_ESTATE) return rc;
    }
    // Small 2-D / 1-D problems whose state is L2-resident (BASELINE configs 1, 2): one persistent cooperative launch for all n
    // iterations instead of n launches ("pd_persist" tuning knob: 0 auto, 1 whenever possible, 2 never).
    if (n >= 2 && gv.comp_y < 0 && vec_ok && !pl->link_on && !pl->halo_above && !pl->halo_below && ctx->pd_persist != 2 &&
        (ctx->pd_persist == 1 || pl->bytes <= ((size_t)96 << 20))) {
        int rc = gv.dtype == NSOL_F32 ? pd_launch_iteration<float, 4>(pl, st, 0, n) : pd_launch_iteration<double, 2>(pl, st, 0, n);
        if (rc == NSOL_OK) return NSOL_OK;
        if (rc != NSOL_ESTATE) return rc;      // NSOL_ESTATE: cooperative launch unavailable -> one launch per iteration
    }
    for (int i = 0; i < n; ++i) {
        int rc;
        if (gv.dtype == NSOL_F32) rc = vec_ok ? pd_launch_iteration<float, 4>(pl, st) : pd_launch_iteration<float, 1>(pl, st);
        else rc = vec_ok ? pd_launch_iteration<double, 2>(pl, st) : pd_launch_iteration<double, 1>(pl, st);
        if (rc != NSOL_OK) return rc;
    }
    return NSOL_OK;
}

// Split iteration for overlapping the z-slab halo exchange with compute: part 1 computes the two
// boundary chunks (whose first/last planes the neighbours need; fetch them with
// nsol_pd_plan_boundary_planes(next=1)), part 2 the interior chunks and advances the plan.
extern "C" int nsol_pd_plan_iterate_part(nsol_pd_plan *pl, int part, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!pl->ready) return nsol_fail(ctx, NSOL_ESTATE, "pd iterate: plan has not been reset with an observation");
    if (part != 1 && part != 2) return nsol_fail(ctx, NSOL_EINVAL, "pd iterate_part: part must be 1 (boundary) or 2 (interior)");
    NSOL_CHECK(nsol_bind_device(ctx));
    NSOL_CHECK(pd_ensure_schedule(pl, pl->it + 1));
    const GridView &gv = pl->gv;
    cudaStream_t st = (cudaStream_t)s;
    const int vecw = gv.dtype == NSOL_F32 ? 4 : 2;
    const bool vec_ok = (gv.nx % vecw) == 0;
    if (gv.dtype == NSOL_F32) return vec_ok ? pd_launch_iteration<float, 4>(pl, st, part) : pd_launch_iteration<float, 1>(pl, st, part);
    return vec_ok ? pd_launch_iteration<double, 2>(pl, st, part) : pd_launch_iteration<double, 1>(pl, st, part);
}

// number of z-chunks an iteration of this plan is cut into (>= 3 is needed for the split form)
extern "C" int nsol_pd_plan_chunks(nsol_pd_plan *pl) {
    if (!pl) return 0;
    const GridView &gv = pl->gv;
    const int vecw = gv.dtype == NSOL_F32 ? 4 : 2;
    int ty, zc;
    pd_tiling(pl->ctx, gv, (gv.nx % vecw) == 0 ? vecw : 1, &ty, &zc, pl->link_on);
    return (gv.nz + zc - 1) / zc;
}

extern "C" int nsol_pd_plan_x_dev(nsol_pd_plan *pl, const void **x_dev) {
    if (!pl || !x_dev) return NSOL_EINVAL;
    *x_dev = pl->x;
    return NSOL_OK;
}

extern "C" int nsol_pd_plan_get_x_dev(nsol_pd_plan *pl, int dtype_out, void *out_dev, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    if (!pl->ready) return nsol_fail(pl->ctx, NSOL_ESTATE, "pd get_x: plan has not been reset");
    NSOL_CHECK(nsol_bind_device(pl->ctx));
    // get_x(): x * x_scale (nsol/solver.py:117-118)
    return nsol_scale_convert(pl->ctx, pl->gv.n * pl->gv.batch, pl->gv.dtype, pl->x, dtype_out, out_dev, pl->desc.x_scale, 0, s);
}

extern "C" int nsol_pd_plan_get_x_host(nsol_pd_plan *pl, double *x_host, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!x_host) return nsol_fail(ctx, NSOL_EINVAL, "pd get_x: output is NULL");
    if (!pl->ready) return nsol_fail(ctx, NSOL_ESTATE, "pd get_x: plan has not been reset");
    NSOL_CHECK(nsol_bind_device(ctx));
    const size_t nv = (size_t)pl->gv.n * pl->gv.batch;
    NSOL_CHECK(pd_ensure_stage(pl, nv * sizeof(double)));
    cudaStream_t st = (cudaStream_t)s;
    NSOL_CHECK(nsol_scale_convert(ctx, (int64_t)nv, pl->gv.dtype, pl->x, NSOL_F64, pl->stage, pl->desc.x_scale, 0, s));
    NSOL_CUDA(ctx, cudaMemcpyAsync(x_host, pl->stage, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    NSOL_CUDA(ctx, cudaStreamSynchronize(st));
    if (pl->link_on) NSOL_CHECK(nsol_pd_plan_link_status(pl, s));
    return NSOL_OK;
}

// ---- pipelined host solve --------------------------------------------------------------------------------------------
// reset of the planes of one transfer group from the float64 staging copy of the observation (and of x0 when it is a different
// array): b' = b / b_scale, x = xbar = x0 / x0_scale, p = 0 -- the same IEEE operations as pd_reset_common, one pass.
template <typename T>
__global__ void pd_reset_range_kernel(long long n, const double *sb, const double *sx, double b_scale, double x0_scale,
                                      T *b, T *x, T *xbar, T *p0, T *p1, T *p2) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double vb = sb[i], vx = sx[i];
        b[i] = (T)(vb / b_scale);
        const T v = (T)(vx / x0_scale);
        x[i] = v;
        xbar[i] = v;
        p0[i] = T(0);
        if (p1) p1[i] = T(0);
        if (p2) p2[i] = T(0);
    }
}

static bool pd_host_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

template <typename T>
static int pd_launch_range(nsol_pd_plan *pl, cudaStream_t st, int c0, int c1, int it, int zc = 0) {
    const PdRange rg = {c0, c1, it, zc};
    const int vecw = sizeof(T) == 4 ? 4 : 2;
    if ((pl->gv.nx % vecw) == 0) {
        if (sizeof(T) == 4) return pd_launch_iteration<float, 4>(pl, st, 0, 0, &rg);
        return pd_launch_iteration<double, 2>(pl, st, 0, 0, &rg);
    }
    if (sizeof(T) == 4) return pd_launch_iteration<float, 1>(pl, st, 0, 0, &rg);
    return pd_launch_iteration<double, 1>(pl, st, 0, 0, &rg);
}

// One whole solve from / to host memory: x_host = x_scale * PD_iterations(b_host, x0_host).  For a large single volume with
// page-locked host buffers the transfers are cut into groups of z-planes and overlap the iterations (DESIGN.md 3.1 "pipelined
// solve"): group j is uploaded on a copy stream while the groups below it advance on a wavefront -- iteration i of group c
// needs iteration i - 1 of groups c - 1, c, c + 1 only, so when group j lands, groups j-1, j-2, ..., j-D run iterations
// 0, 1, ..., D-1; after the last upload the volume is brought to D iterations everywhere, iterations D ... n-D'-1 are whole-volume
// launches, and the mirror image at the end lets group 0 finish first and go back to the host while the groups above it catch up.
// The kernels, the chunk geometry and therefore every bit of the result are those of nsol_pd_plan_iterate.
void nsol_preload_scale_convert();      // capi.cu

// load every kernel the pipelined solve of this plan is going to launch (see g_pd_query_only)
template <typename T>
static int pd_preload(nsol_pd_plan *pl, cudaStream_t st) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, pd_reset_range_kernel<T>);
    cudaFuncGetAttributes(&fa, pd_link_publish_kernel);
    nsol_preload_scale_convert();
    g_pd_query_only = true;
    const int rc = pd_launch_range<T>(pl, st, 0, 1, 0);
    g_pd_query_only = false;
    cudaGetLastError();
    return rc;
}

template <typename T>
static int pd_solve_pipelined(nsol_pd_plan *pl, const double *b_host, const double *x0_host, int iterations, double *x_host,
                              cudaStream_t st, int planes, int depth) {
    nsol_ctx *ctx = pl->ctx;
    const GridView &gv = pl->gv;
    const int vecw = sizeof(T) == 4 ? 4 : 2;
    int ty, zc;
    pd_tiling(ctx, gv, (gv.nx % vecw) == 0 ? vecw : 1, &ty, &zc, pl->link_on);
    // transfer groups: 16 planes, 8 or 4 for a thin slab so that it still has ~16 groups -- the wavefront gains one iteration of
    // skew per group, and on a multi-GPU box whose host feeds every GPU at a fraction of its link rate the transfers are worth
    // dozens of iterations (8 GPUs, 64-plane slabs: 8.7 ms per direction = 33 iterations).  Measured on 2 x 64 planes
    // (profiles/r2_thin_slabs.md): 4 planes 29.0 ms per solve, 16 planes 29.8, 2 planes 30.7-35.0 (launches too short).
    if (planes <= 0) {
        planes = 16;
        while (planes > 4 && gv.nz / planes < 16) planes /= 2;
    }
    if (depth <= 0) depth = planes < 8 ? 12 : 10;
    const int zr = std::max(1, std::min(zc, planes));               // planes per z-chunk of the chunk-range launches
    const int nchunks = (gv.nz + zr - 1) / zr;
    const int gch = planes / zr > 0 ? planes / zr : 1;              // z-chunks per transfer group
    const int ng = (nchunks + gch - 1) / gch;
    const long long plane = (long long)gv.nx * gv.ny;
    const bool same = (x0_host == nullptr) || (x0_host == b_host);
    const size_t nv = (size_t)gv.n;
    NSOL_CHECK(pd_ensure_stage(pl, (same ? 1 : 2) * nv * sizeof(double)));
    NSOL_CHECK(pd_ensure_schedule(pl, iterations));
    if (!pl->pipe_up) {
        NSOL_CUDA(ctx, cudaStreamCreateWithFlags(&pl->pipe_up, cudaStreamNonBlocking));
        NSOL_CUDA(ctx, cudaStreamCreateWithFlags(&pl->pipe_dn, cudaStreamNonBlocking));
        NSOL_CUDA(ctx, cudaEventCreateWithFlags(&pl->pipe_fork, cudaEventDisableTiming));
    }
    for (auto *v : {&pl->pipe_ev_up, &pl->pipe_ev_x, &pl->pipe_ev_dn})
        while ((int)v->size() < ng) {
            cudaEvent_t e;
            NSOL_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            v->push_back(e);
        }
    double *sb = (double *)pl->stage, *sx = same ? sb : sb + nv;
    // Groups are numbered in ARRIVAL order: arrival index c is the spatial group c when the transfers run bottom-up and
    // ng - 1 - c when they run top-down.  Neighbouring z-slabs of a linked decomposition run in opposite directions, so the two
    // groups that face each other across a slab boundary are both first or both last to arrive -- the wavefront then continues
    // through the boundary (the boundary groups' launches wait for / publish the halo planes by iteration number).
    const bool down = pl->link_on && pl->pipe_dir < 0;
    auto spatial = [&](int c) { return down ? ng - 1 - c : c; };
    auto g_lo = [&](int g) { return (long long)std::min(g * gch * zr, gv.nz) * plane; };     // first voxel of spatial group g
    auto chunk_lo = [&](int g) { return std::min(g * gch, nchunks); };
    const int d_up = std::min(depth, iterations / 2), d_dn = std::min(depth, iterations - d_up);
    std::vector<int> done(ng, -1);                                   // by arrival index; -1: not yet on the device
    auto advance = [&](int c) -> int {
        // the wavefront rule: both neighbours hold the state this iteration reads
        if (done[c] < 0 || (c > 0 && done[c - 1] < done[c]) || (c + 1 < ng && done[c + 1] < done[c]))
            return nsol_fail(ctx, NSOL_ESTATE, "pd pipelined solve: wavefront order violated at group %d", c);
        const int g = spatial(c);
        NSOL_CHECK(pd_launch_range<T>(pl, st, chunk_lo(g), chunk_lo(g + 1), done[c], zr));
        done[c] += 1;
        return NSOL_OK;
    };
    // a merged launch of the arrival groups [0, top], all at iteration `it`
    auto advance_first = [&](int top, int it) -> int {
        const int ga = spatial(0), gb = spatial(top);
        return pd_launch_range<T>(pl, st, chunk_lo(std::min(ga, gb)), chunk_lo(std::max(ga, gb) + 1), it, zr);
    };
    NSOL_CHECK(pd_preload<T>(pl, st));
    NSOL_CHECK(pd_push_join(pl, st));           // a queued publish kernel may still read the state this solve overwrites
    pl->cur = 0;
    pl->it = 0;
    pl->pipe_g0 = pl->link_pub;          // generation of this solve's start state
    // ---- uploads, all queued on the copy stream behind whatever the caller's stream holds
    NSOL_CUDA(ctx, cudaEventRecord(pl->pipe_fork, st));
    NSOL_CUDA(ctx, cudaStreamWaitEvent(pl->pipe_up, pl->pipe_fork, 0));
    for (int c = 0; c < ng; ++c) {
        const int g = spatial(c);
        const long long lo = g_lo(g), cnt = g_lo(g + 1) - lo;
        NSOL_CUDA(ctx, cudaMemcpyAsync(sb + lo, b_host + lo, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, pl->pipe_up));
        if (!same) NSOL_CUDA(ctx, cudaMemcpyAsync(sx + lo, x0_host + lo, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, pl->pipe_up));
        NSOL_CUDA(ctx, cudaEventRecord(pl->pipe_ev_up[c], pl->pipe_up));
    }
    // ---- arrival phase: group j lands -> reset it, groups j-1 ... j-D advance one iteration each (newest first)
    for (int j = 0; j < ng + d_up; ++j) {
        if (j < ng) {
            const int g = spatial(j);
            const long long lo = g_lo(g), cnt = g_lo(g + 1) - lo;
            NSOL_CUDA(ctx, cudaStreamWaitEvent(st, pl->pipe_ev_up[j], 0));
            const int threads = 256;
            const long long want = (cnt + threads - 1) / threads;
            const int blocks = (int)std::min<long long>(std::max<long long>(want, 1), (long long)ctx->sm_count * 16);
            T *p1 = gv.dim > 1 ? (T *)pl->p[0][1] + lo : nullptr, *p2 = gv.dim > 2 ? (T *)pl->p[0][2] + lo : nullptr;
            pd_reset_range_kernel<T><<<blocks, threads, 0, st>>>(cnt, sb + lo, sx + lo, pl->desc.b_scale, pl->desc.x0_scale, (T *)pl->b + lo,
                                                                (T *)pl->x + lo, (T *)pl->xbar[0] + lo, (T *)pl->p[0][0] + lo, p1, p2);
            NSOL_LAUNCH_CHECK(ctx);
            done[j] = 0;
            if (pl->link_on) {
                // the start state of a boundary group goes to the neighbour as soon as it exists
                if (g == 0) NSOL_CHECK(pd_link_publish(pl, st, 1));
                if (g == ng - 1) NSOL_CHECK(pd_link_publish(pl, st, 2));
            }
        }
        for (int c = std::min(j - 1, ng - 1); c >= std::max(0, j - d_up); --c) NSOL_CHECK(advance(c));
    }
    // ---- whole-volume iterations d_up ... iterations - d_dn - 1
    pl->cur = d_up & 1;
    pl->it = d_up;
    pl->ready = true;
    if (pl->link_on) {
        pl->link_pub = pl->pipe_g0 + 1u + (unsigned)d_up;
        pl->link_fresh = true;
    }
    for (int i = d_up; i < iterations - d_dn; ++i) {
        if (sizeof(T) == 4) NSOL_CHECK(((gv.nx % 4) == 0 ? pd_launch_iteration<float, 4>(pl, st) : pd_launch_iteration<float, 1>(pl, st)));
        else NSOL_CHECK(((gv.nx % 2) == 0 ? pd_launch_iteration<double, 2>(pl, st) : pd_launch_iteration<double, 1>(pl, st)));
    }
    for (int c = 0; c < ng; ++c) done[c] = iterations - d_dn;
    // ---- ramp: round t advances the arrival groups [0, d_dn - t] (one launch, they are at the same iteration) -> group c is c
    // iterations short of the end
    for (int t = 1; t <= d_dn; ++t) {
        const int top = std::min(d_dn - t, ng - 1);
        NSOL_CHECK(advance_first(top, iterations - d_dn + t - 1));
        for (int c = 0; c <= top; ++c) done[c] += 1;
    }
    // ---- departure phase: step j finishes group j (farthest group first), converts it and sends it home on the copy stream
    for (int j = 0; j < ng; ++j) {
        for (int c = std::min(j + d_dn - 1, ng - 1); c >= j; --c)
            if (done[c] < iterations && (c + 1 >= ng || done[c + 1] >= done[c])) NSOL_CHECK(advance(c));
        if (done[j] != iterations) return nsol_fail(ctx, NSOL_ESTATE, "pd pipelined solve: group %d stopped at iteration %d", j, done[j]);
        const int g = spatial(j);
        const long long lo = g_lo(g), cnt = g_lo(g + 1) - lo;
        NSOL_CHECK(nsol_scale_convert(ctx, cnt, gv.dtype, (const T *)pl->x + lo, NSOL_F64, sb + lo, pl->desc.x_scale, 0, (nsol_stream)st));
        NSOL_CUDA(ctx, cudaEventRecord(pl->pipe_ev_x[j], st));
        NSOL_CUDA(ctx, cudaStreamWaitEvent(pl->pipe_dn, pl->pipe_ev_x[j], 0));
        NSOL_CUDA(ctx, cudaMemcpyAsync(x_host + lo, sb + lo, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, pl->pipe_dn));
    }
    NSOL_CUDA(ctx, cudaEventRecord(pl->pipe_ev_dn[0], pl->pipe_dn));
    NSOL_CUDA(ctx, cudaStreamWaitEvent(st, pl->pipe_ev_dn[0], 0));
    pl->cur = iterations & 1;
    pl->it = iterations;
    if (pl->link_on) {
        pl->link_pub = pl->pipe_g0 + 1u + (unsigned)iterations;
        pl->link_fresh = true;
    }
    pl->pipe_groups_last = ng;
    pl->pipe_depth_last = d_up;
    NSOL_CUDA(ctx, cudaStreamSynchronize(st));
    if (pl->link_on) NSOL_CHECK(nsol_pd_plan_link_status(pl, (nsol_stream)st));
    return NSOL_OK;
}

extern "C" int nsol_pd_plan_solve_host(nsol_pd_plan *pl, const double *b_host, const double *x0_host, int iterations,
                                       double *x_host, nsol_stream s) {
    if (!pl) return NSOL_EINVAL;
    nsol_ctx *ctx = pl->ctx;
    if (!b_host || !x_host) return nsol_fail(ctx, NSOL_EINVAL, "pd solve: b or x is NULL");
    if (iterations < 0) return nsol_fail(ctx, NSOL_EINVAL, "pd solve: iterations must be >= 0");
    NSOL_CHECK(nsol_bind_device(ctx));
    const GridView &gv = pl->gv;
    pl->pipe_groups_last = 0;
    // z-slabs: only with the in-kernel halo exchange (caller-refreshed halo buffers belong to whole-slab iterations)
    const bool possible = gv.comp_z >= 0 && gv.batch == 1 && (pl->link_on || (!pl->halo_above && !pl->halo_below)) && ctx->pd_pipe != 2;
    if (possible && (ctx->pd_pipe == 1 || ((size_t)gv.n * sizeof(double) >= ((size_t)64 << 20) && pd_host_pinned(b_host) &&
                                           pd_host_pinned(x_host) && (!x0_host || x0_host == b_host || pd_host_pinned(x0_host))))) {
        const int planes = ctx->pd_pipe_planes, depth = ctx->pd_pipe_depth;     // 0: chosen from the slab height
        if (gv.dtype == NSOL_F32) return pd_solve_pipelined<float>(pl, b_host, x0_host, iterations, x_host, (cudaStream_t)s, planes, depth);
        return pd_solve_pipelined<double>(pl, b_host, x0_host, iterations, x_host, (cudaStream_t)s, planes, depth);
    }
    NSOL_CHECK(nsol_pd_plan_reset_host(pl, b_host, x0_host, s));
    NSOL_CHECK(nsol_pd_plan_iterate(pl, iterations, s));
    return nsol_pd_plan_get_x_host(pl, x_host, s);
}

// Direction in which the transfer groups of nsol_pd_plan_solve_host travel through a linked z-slab: +1 bottom-up, -1 top-down.
// Neighbouring slabs must run in opposite directions (rank parity) for the wavefront to continue through their boundary.
extern "C" int nsol_pd_plan_set_pipe_direction(nsol_pd_plan *pl, int direction) {
    if (!pl) return NSOL_EINVAL;
    pl->pipe_dir = direction < 0 ? -1 : 1;
    return NSOL_OK;
}

// transfer groups and wavefront depth of the last nsol_pd_plan_solve_host (0 groups: the plain upload / iterate / download sequence)
extern "C" int nsol_pd_plan_solve_info(const nsol_pd_plan *pl, int *groups_out, int *depth_out) {
    if (!pl) return NSOL_EINVAL;
    if (groups_out) *groups_out = pl->pipe_groups_last;
    if (depth_out) *depth_out = pl->pipe_depth_last;
    return NSOL_OK;
}

extern "C" int nsol_pd_run_host(nsol_ctx *ctx, const nsol_pd_desc *desc, int iterations, const double *b_host,
                                const double *x0_host, double *x_host, double *iterates_host, nsol_stream s) {
    if (!ctx) return NSOL_EINVAL;
    if (iterations < 0) return nsol_fail(ctx, NSOL_EINVAL, "pd run: iterations must be >= 0");
    nsol_pd_plan *pl = nullptr;
    NSOL_CHECK(nsol_pd_plan_create(ctx, desc, &pl));
    int rc = nsol_pd_plan_reset_host(pl, b_host, x0_host, s);
    const size_t nv = (size_t)pl->gv.n * pl->gv.batch;
    if (rc == NSOL_OK && iterates_host) {
        rc = nsol_pd_plan_get_x_host(pl, iterates_host, s);
        for (int i = 0; rc == NSOL_OK && i < iterations; ++i) {
            rc = nsol_pd_plan_iterate(pl, 1, s);
            if (rc == NSOL_OK) rc = nsol_pd_plan_get_x_host(pl, iterates_host + (size_t)(i + 1) * nv, s);
        }
    } else if (rc == NSOL_OK) {
        rc = nsol_pd_plan_iterate(pl, iterations, s);
    }
    if (rc == NSOL_OK) rc = nsol_pd_plan_get_x_host(pl, x_host, s);
    nsol_pd_plan_destroy(pl);
    return rc;
}
