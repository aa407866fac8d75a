// lsmr_fused3d.cuh -- fused 3-D forward / adjoint kernels of the LSMR iteration on [A; sqrt(alpha) grad]:
// the whole separable periodic blur (x, y AND z) is evaluated inside the consumer, no pass kernel, no scratch array.
//
//   forward   u    <- (u * inv_beta) * (-alpha) + [A v; sqrt_alpha grad v],  v = vhat * inv_alpha     (scipy lsmr.py:336-338;
//                                                       operator nsol/tikhonov_linear_solver.py:258-264)
//   adjoint   vhat <- (vhat * inv_alpha) * (-beta) + (A^T u0 + sqrt_alpha grad^T u1..3) * inv_beta     (lsmr.py:342-344;
//                                                       operator nsol/tikhonov_linear_solver.py:266-274)
//
// With separate pass kernels along y and z the 3-D path moved 30 words per voxel and inner iteration against the
// algorithmic 22 (9 forward + 6 adjoint + 7 update) and ran at 0.32 of the HBM roofline
// (profiles/r1_lsmr_v2.md).  Here a CTA of TY warps owns a tile of TY rows x W = 32 * VEC columns and marches
// through a chunk of z-planes:
//   * every input plane of the blurred array (vhat resp. u block 0) is staged ONCE as a raw tile of TY + 2R rows x
//     W + 2A columns (periodic halos) in a ring of shared-memory stages, filled P planes ahead with cp.async (LDGSTS:
//     bytes in flight without registers);
//   * the blur along y reads 2R + 1 rows of the newest raw tile, the blur along x goes through a warp-private
//     shared-memory row (own columns + halo vectors), and the x/y-blurred plane enters a register ring of 2R + 1
//     planes whose dot product with the z taps is the blurred value of the plane R steps behind;
//   * forward: the gradient of plane z needs the raw vhat planes z and z + 1 -- they are still in the stage ring
//     (S = R + P + 1 stages), so vhat is read from DRAM once; the four u blocks of the plane are staged P planes
//     ahead in an operand ring and written back in place;
//   * adjoint: u1 (with the column to the left), u2 (with the row below), u3 and vhat of the plane are staged the
//     same way, u3 of the previous plane is carried in registers.
// DRAM traffic = the algorithmic 9 / 6 words per voxel plus the 2R / zc warm-up planes of a chunk; halo rows and
// columns of neighbouring tiles are L2 hits.  The blur is accumulated with fused multiply-adds in the order y, x, z
// (the pass kernels go z, y, x without contraction): same taps, different rounding -- compared at 1e-11 in
// tests/test_gpu_parity.py.  Zero (Dirichlet) boundary of the gradient, periodic boundary of the blur, as in the
// reference (nsol/linear_operators.py:98-106 mode="constant", :60-68 mode="wrap").
#pragma once

#define F3_TY 8
#define F3_P 2

template <typename T, int VEC, int R>
struct F3 {
    static constexpr int A = (R + VEC - 1) / VEC * VEC;     // halo columns, whole vectors
    static constexpr int HV = A / VEC;                      // halo vectors per side
    static constexpr int W = 32 * VEC;
    static constexpr int LEN = W + 2 * A;
    static constexpr int ROWS = F3_TY + 2 * R;
    static constexpr int S = R + F3_P + 1;                  // raw stages: planes j - R .. j + P
    static constexpr int OS = F3_P + 1;                     // operand stages
    static constexpr int RAW = ROWS * LEN;
    static constexpr int OPF = 4 * F3_TY * W;               // forward: u0..u3
    static constexpr int U1LEN = W + VEC;                   // adjoint: u1 row with one vector to the left
    static constexpr int OPA = F3_TY * U1LEN + (F3_TY + 1) * W + 2 * F3_TY * W;   // u1 | u2 (+ row below) | u3 | vhat
    static size_t smem(bool fwd) { return sizeof(T) * ((size_t)S * RAW + (size_t)F3_TY * LEN + (size_t)OS * (fwd ? OPF : OPA)); }
};

struct F3Geom {
    int nx, ny, nz, zc;
    long long n;
    // z-slab decomposition (nsol_lsmr_plan_slab): planes below 0 / above nz - 1 of the blurred array come from the halo
    // buffers filled by the neighbour exchange (ring: periodic blur) instead of the in-volume wrap; the gradient has a
    // neighbour below / above or the zero boundary at the global ends
    int slab, grad_lo, grad_hi, ghost;
};

__device__ __forceinline__ void f3_cp16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void f3_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void f3_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ double f3_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float f3_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// sum over a (32, F3_TY) block; result valid in thread (0, 0).  (block_sum of lsmr_kernels.cu assumes a 1-D block.)
__device__ __forceinline__ double f3_block_sum(double v) {
    __shared__ double warp_part[F3_TY];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) warp_part[threadIdx.y] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
#pragma unroll
        for (int w = 0; w < F3_TY; ++w) r += warp_part[w];      // fixed order: deterministic
    }
    return r;
}

__device__ __forceinline__ int f3_mod(int v, int n) {
    v %= n;
    return v < 0 ? v + n : v;
}

template <typename T, int R, int VEC, bool FWD>
__global__ void __launch_bounds__(32 * F3_TY, 2) fused3d_kernel(F3Geom g, T wx, T wy, T wz, const LsmrScalars *__restrict__ S, TapsR<T, R> tx,
                                                                TapsR<T, R> ty_, TapsR<T, R> tz, const T *__restrict__ blur_in, T *__restrict__ u,
                                                                T *__restrict__ vhat, double *__restrict__ part, int first,
                                                                const T *__restrict__ halo_lo, const T *__restrict__ halo_hi,
                                                                const T *__restrict__ uz_lo) {
    using V = Vec<T, VEC>;
    using L = F3<T, VEC, R>;
    extern __shared__ __align__(16) unsigned char f3_smem[];
    if (S->done) return;
    T *s_raw = reinterpret_cast<T *>(f3_smem);              // [S][ROWS][LEN]
    T *s_yb = s_raw + L::S * L::RAW;                        // [TY][LEN]
    T *s_op = s_yb + F3_TY * L::LEN;                        // [OS][OPF | OPA]
    constexpr int OPSZ = FWD ? L::OPF : L::OPA;

    const int lane = threadIdx.x, ty = threadIdx.y;
    const int x0t = (int)blockIdx.x * L::W;
    const int y0 = (int)blockIdx.y * F3_TY;
    const int z0 = (int)blockIdx.z * g.zc;
    const int z1 = min(g.nz, z0 + g.zc);
    const int x = x0t + lane * VEC;
    const int y = y0 + ty;
    const bool col_in = x < g.nx;
    const bool active = col_in && y < g.ny;
    const int wcols = min(L::W, g.nx - x0t);                // columns of the tile inside the volume
    const long long plane = (long long)g.nx * g.ny;
    const T inv_alpha = (T)S->inv_alpha, inv_beta = (T)S->inv_beta, sa = (T)S->sqrt_alpha;
    const T mscale = FWD ? (T)(-S->alpha) : (T)(-S->beta);

    // ---- what this thread copies for every raw plane: its VEC columns (+ one halo vector for lanes < 2 HV) of the
    //      tile rows ty and ty + TY (tile row r <-> volume row y0 - R + r, periodic)
    const bool has_halo = R > 0 && lane < 2 * L::HV;
    int hcol = lane < L::HV ? x0t - L::A + lane * VEC : x0t + wcols + (lane - L::HV) * VEC;
    hcol = f3_mod(hcol, g.nx);
    const int hslot = lane < L::HV ? lane * VEC : L::A + wcols + (lane - L::HV) * VEC;
    constexpr int NR = (L::ROWS + F3_TY - 1) / F3_TY;       // tile rows per thread: ty, ty + TY, (ty + 2 TY)
    long long rowoff[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) rowoff[q] = (long long)f3_mod(y0 - R + min(ty + q * F3_TY, L::ROWS - 1), g.ny) * g.nx;

    auto issue_raw = [&](int j) {        // raw plane of ring index j: volume plane z0 - R + j (periodic)
        int zi = z0 - R + j;
        const T *src;
        if (g.slab) {            // planes beyond the slab: the neighbours' planes in the halo buffers
            src = zi < 0 ? halo_lo + (long long)(zi + g.ghost) * plane : (zi >= g.nz ? halo_hi + (long long)(zi - g.nz) * plane : blur_in + (long long)zi * plane);
        } else {
            zi += zi < 0 ? g.nz : 0;
            zi -= zi >= g.nz ? g.nz : 0;
            src = blur_in + (long long)zi * plane;
        }
        T *st = s_raw + (j % L::S) * L::RAW;
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const int r = ty + q * F3_TY;
            if (r < L::ROWS) {
                if (col_in) f3_cp16(st + r * L::LEN + L::A + lane * VEC, src + rowoff[q] + x);
                if (has_halo) f3_cp16(st + r * L::LEN + hslot, src + rowoff[q] + hcol);
            }
        }
    };
    const T *u0 = u, *u1 = u + g.n, *u2 = u + 2 * g.n, *u3 = u + 3 * g.n;
    auto issue_ops = [&](int j) {        // operands of the output plane of step j: z = z0 + j - 2R
        const int z = z0 + j - 2 * R;
        if (z < z0 || z >= z1 || !active) return;
        T *st = s_op + (j % L::OS) * OPSZ;
        const long long i = (long long)z * plane + (long long)y * g.nx + x;
        if (FWD) {
            f3_cp16(st + (0 * F3_TY + ty) * L::W + lane * VEC, u0 + i);
            f3_cp16(st + (1 * F3_TY + ty) * L::W + lane * VEC, u1 + i);
            f3_cp16(st + (2 * F3_TY + ty) * L::W + lane * VEC, u2 + i);
            f3_cp16(st + (3 * F3_TY + ty) * L::W + lane * VEC, u3 + i);
        } else {
            T *s1 = st, *s2 = st + F3_TY * L::U1LEN, *s3 = s2 + (F3_TY + 1) * L::W, *sv = s3 + F3_TY * L::W;
            f3_cp16(s1 + ty * L::U1LEN + VEC + lane * VEC, u1 + i);
            if (lane == 0 && x > 0) f3_cp16(s1 + ty * L::U1LEN, u1 + i - VEC);        // the vector left of the tile row
            f3_cp16(s2 + (ty + 1) * L::W + lane * VEC, u2 + i);
            if (ty == 0 && y > 0) f3_cp16(s2 + lane * VEC, u2 + i - g.nx);            // the row below the tile
            f3_cp16(s3 + ty * L::W + lane * VEC, u3 + i);
            if (!first) f3_cp16(sv + ty * L::W + lane * VEC, vhat + i);
        }
    };

    const int steps = (z1 - z0) + 2 * R;
    V ring[2 * R + 1];
#pragma unroll
    for (int i = 0; i <= 2 * R; ++i) ring[i] = vec_zero<T, VEC>();
    V u3_prev = vec_zero<T, VEC>();                          // adjoint: u3 of plane z - 1 (zero below the first plane)
    if (!FWD && active && z0 > 0) u3_prev = vec_load<T, VEC>(u3 + (long long)(z0 - 1) * plane + (long long)y * g.nx + x);
    if (!FWD && active && z0 == 0 && g.slab && g.grad_lo) u3_prev = vec_load<T, VEC>(uz_lo + (long long)y * g.nx + x);   // lower neighbour's last u_z plane
    double acc = 0.0;

    // ---- prologue: planes 0 .. P-1 (one group each) -----------------------------------------------------------
#pragma unroll
    for (int j = 0; j < F3_P; ++j) {
        if (j < steps) {
            issue_raw(j);
            issue_ops(j);
        }
        f3_commit();
    }

    for (int j = 0; j < steps; ++j) {
        f3_wait<F3_P - 1>();             // the group of plane j (issued P steps ago) has landed for this thread
        __syncthreads();                 // ... and for everybody; all threads are done with step j - 1
        if (j + F3_P < steps) {
            issue_raw(j + F3_P);         // overwrites the stage of plane j - R - 1: last read in step j - 1
            issue_ops(j + F3_P);
        }
        f3_commit();

        // ---- blur along y of the newest raw plane, own columns and the halo vector, into the warp's row ----------
        const T *raw = s_raw + (j % L::S) * L::RAW;
        T *yb = s_yb + ty * L::LEN;
        {
            V a = vec_zero<T, VEC>(), h = vec_zero<T, VEC>();
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {               // volume row y - (k - R) = tile row ty + 2R - k
                const T *row = raw + (ty + 2 * R - k) * L::LEN;
                if (col_in) {
                    const V w = vec_load<T, VEC>(row + L::A + lane * VEC);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) a.v[v] = f3_fma(ty_.t[k], w.v[v], a.v[v]);
                }
                if (has_halo) {
                    const V w = vec_load<T, VEC>(row + hslot);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) h.v[v] = f3_fma(ty_.t[k], w.v[v], h.v[v]);
                }
            }
            if (col_in) vec_store<T, VEC>(yb + L::A + lane * VEC, a);
            if (has_halo) vec_store<T, VEC>(yb + hslot, h);
        }
        __syncwarp();
        // ---- blur along x from the warp's row ---------------------------------------------------------------
        V hx = vec_zero<T, VEC>();
        if (col_in) {
            T val[VEC + 2 * L::A];
#pragma unroll
            for (int q = 0; q < (VEC + 2 * L::A) / VEC; ++q) {
                const V w = vec_load<T, VEC>(yb + lane * VEC + q * VEC);
#pragma unroll
                for (int v = 0; v < VEC; ++v) val[q * VEC + v] = w.v[v];
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                T s = T(0);
#pragma unroll
                for (int k = 0; k <= 2 * R; ++k) s = f3_fma(tx.t[k], val[L::A + v + R - k], s);     // column (x + v) - (k - R)
                hx.v[v] = s;
            }
        }
        __syncwarp();                    // the row is rewritten in the next step
#pragma unroll
        for (int i = 0; i < 2 * R; ++i) ring[i] = ring[i + 1];
        ring[2 * R] = hx;                // ring[i] = x/y-blurred plane z - R + i
        if (j < 2 * R) continue;

        // ---- output plane z ---------------------------------------------------------------------------------------
        const int z = z0 + j - 2 * R;
        const long long i0 = (long long)z * plane + (long long)y * g.nx + x;
        const T *op = s_op + (j % L::OS) * OPSZ;
        if (active) {
            V av;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                T s = T(0);
#pragma unroll
                for (int k = 0; k <= 2 * R; ++k) s = f3_fma(tz.t[k], ring[2 * R - k].v[v], s);      // plane z - (k - R)
                av.v[v] = s;
            }
            if (FWD) {
                // raw vhat of plane z (ring index j - R) and z + 1 (j - R + 1) are still staged
                const T *rz = s_raw + ((j - R) % L::S) * L::RAW + (ty + R) * L::LEN + L::A + lane * VEC;
                const T *rz1 = s_raw + ((j - R + 1) % L::S) * L::RAW + (ty + R) * L::LEN + L::A + lane * VEC;
                const V vr = vec_load<T, VEC>(rz);
                const T right = (x + VEC < g.nx) ? rz[VEC] : T(0);                                  // zero boundary, not the periodic halo
                V vup = vec_zero<T, VEC>(), vz = vec_zero<T, VEC>();
                if (y + 1 < g.ny) vup = vec_load<T, VEC>(rz + L::LEN);
                if (z + 1 < g.nz || (g.slab && g.grad_hi)) vz = vec_load<T, VEC>(rz1);      // (slab: the staged plane nz is the upper neighbour's first plane)
                V a0 = vec_load<T, VEC>(op + (0 * F3_TY + ty) * L::W + lane * VEC);
                V a1 = vec_load<T, VEC>(op + (1 * F3_TY + ty) * L::W + lane * VEC);
                V a2 = vec_load<T, VEC>(op + (2 * F3_TY + ty) * L::W + lane * VEC);
                V a3 = vec_load<T, VEC>(op + (3 * F3_TY + ty) * L::W + lane * VEC);
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
                    const T dy = wy * (vup.v[v] * inv_alpha) + (-wy) * vc.v[v];
                    un = (a2.v[v] * inv_beta) * mscale + sa * dy;
                    a2.v[v] = un;
                    acc += (double)un * (double)un;
                    const T dz = wz * (vz.v[v] * inv_alpha) + (-wz) * vc.v[v];
                    un = (a3.v[v] * inv_beta) * mscale + sa * dz;
                    a3.v[v] = un;
                    acc += (double)un * (double)un;
                }
                vec_store<T, VEC>(u + i0, a0);
                vec_store<T, VEC>(u + g.n + i0, a1);
                vec_store<T, VEC>(u + 2 * g.n + i0, a2);
                vec_store<T, VEC>(u + 3 * g.n + i0, a3);
            } else {
                const T *s1 = op, *s2 = op + F3_TY * L::U1LEN, *s3 = s2 + (F3_TY + 1) * L::W, *sv = s3 + F3_TY * L::W;
                const V b1 = vec_load<T, VEC>(s1 + ty * L::U1LEN + VEC + lane * VEC);
                const T left = (x > 0) ? s1[ty * L::U1LEN + VEC + lane * VEC - 1] : T(0);
                const V b2 = vec_load<T, VEC>(s2 + (ty + 1) * L::W + lane * VEC);
                V b2dn = vec_zero<T, VEC>();
                if (y > 0) b2dn = vec_load<T, VEC>(s2 + ty * L::W + lane * VEC);
                const V b3 = vec_load<T, VEC>(s3 + ty * L::W + lane * VEC);
                V vv = vec_zero<T, VEC>();
                if (!first) vv = vec_load<T, VEC>(sv + ty * L::W + lane * VEC);
                const T lft = left * inv_beta;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    T r = av.v[v] * inv_beta;
                    const T lo = (v == 0) ? lft : b1.v[(v + VEC - 1) % VEC] * inv_beta;
                    T div = wx * lo + (-wx) * (b1.v[v] * inv_beta);
                    div = div + (wy * (b2dn.v[v] * inv_beta) + (-wy) * (b2.v[v] * inv_beta));
                    div = div + (wz * (u3_prev.v[v] * inv_beta) + (-wz) * (b3.v[v] * inv_beta));
                    r = r + sa * div;
                    const T vn = first ? r : (vv.v[v] * inv_alpha) * mscale + r;
                    vv.v[v] = vn;
                    acc += (double)vn * (double)vn;
                }
                vec_store<T, VEC>(vhat + i0, vv);
                u3_prev = b3;
            }
        }
    }
    f3_wait<0>();
    acc = f3_block_sum(acc);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
}

// when the fused 3-D kernels apply: 3-D, A = separable blur with one radius (1..FASTV_MAX_R) on all axes, B = grad, no slab
// mode, every axis longer than the mask; "lsmr_fuse3d" tuning knob: 0 auto (volumes from 16 MB per vector), 1 whenever
// possible, 2 never
static bool fused3d_ok(const nsol_lsmr_plan *pl, int b_op) {
    const GridView &gv = pl->gv;
    if (gv.dim != 3 || pl->desc.a_op != NSOL_A_BLUR || b_op != NSOL_B_GRAD || !fastv_ok(pl)) return false;
    if (pl->ctx->lsmr_fuse3d == 2) return false;
    const int r = pl->desc.radius[0];
    if (r < 1 || r != pl->desc.radius[1] || r != pl->desc.radius[2]) return false;
    if (gv.nz < 2 * r + 1 || gv.ny <= r || gv.nx <= 2 * r) return false;
    if (gv.ny > 65535 * F3_TY) return false;
    if (pl->ctx->lsmr_fuse3d == 1) return true;
    return gv.n * (long long)pl->esz >= (1ll << 24);
}

static int fused3d_planes_per_chunk(const nsol_lsmr_plan *pl, int vec) {
    const GridView &gv = pl->gv;
    const long long tiles = (long long)((gv.nx + 32 * vec - 1) / (32 * vec)) * ((gv.ny + F3_TY - 1) / F3_TY);
    // two CTAs of 256 threads are resident per SM: aim at >= 3 waves, never fewer than 16 planes per chunk (warm-up 2R planes)
    int zc = 64;
    while (zc > 16 && tiles * ((gv.nz + zc - 1) / zc) < (long long)pl->ctx->sm_count * 6) zc /= 2;
    if (pl->ctx->lsmr_fuse3d == 1 && gv.nz < 64) zc = gv.nz < 8 ? gv.nz : 8;     // tests: several chunks on small volumes
    if (zc > gv.nz) zc = gv.nz;
    return zc;
}

template <typename T, int R, bool FWD>
static int fused3d_launch_r(nsol_lsmr_plan *pl, int first, cudaStream_t s, int *nparts) {
    constexpr int VEC = FastvCfg<T>::VEC;
    using L = F3<T, VEC, R>;
    const GridView &gv = pl->gv;
    F3Geom g;
    g.nx = gv.nx;
    g.ny = gv.ny;
    g.nz = gv.nz;
    g.n = gv.n;
    g.zc = fused3d_planes_per_chunk(pl, VEC);
    g.slab = pl->slab ? 1 : 0;
    g.grad_lo = pl->grad_lo;
    g.grad_hi = pl->grad_hi;
    g.ghost = pl->ghost;
    const dim3 grid((gv.nx + L::W - 1) / L::W, (gv.ny + F3_TY - 1) / F3_TY, (gv.nz + g.zc - 1) / g.zc);
    const dim3 block(32, F3_TY, 1);
    const size_t smem = L::smem(FWD);
    static bool configured[64] = {false};          // per device; set once per instantiation (idempotent: a race only repeats the call)
    const int dev = pl->ctx->device & 63;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fused3d_kernel<T, R, VEC, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return nsol_fail(pl->ctx, NSOL_ECUDA, "fused3d: smem opt-in %zu -> %s", smem, cudaGetErrorString(e));
        configured[dev] = true;
    }
    // numpy axis 0 = z, 1 = y, 2 = x; derivative component k acts on axis 2 - k
    fused3d_kernel<T, R, VEC, FWD><<<grid, block, smem, s>>>(g, (T)gv.w[0], (T)gv.w[1], (T)gv.w[2], pl->S, lsq_taps_r<T, R>(pl, 2), lsq_taps_r<T, R>(pl, 1),
                                                            lsq_taps_r<T, R>(pl, 0), FWD ? (const T *)pl->v : (const T *)pl->u, (T *)pl->u,
                                                            (T *)pl->v, pl->part, first, (const T *)(FWD ? pl->halo_v_lo : pl->halo_u_lo),
                                                            (const T *)(FWD ? pl->halo_v_hi : pl->halo_u_hi), (const T *)pl->halo_uz_lo);
    *nparts = (int)(grid.x * grid.y * grid.z);
    NSOL_LAUNCH_CHECK(pl->ctx);
    return NSOL_OK;
}

template <typename T>
static int fused3d_launch(nsol_lsmr_plan *pl, bool forward, int first, cudaStream_t s, int *nparts) {
    const int r = pl->desc.radius[0];
#define F3_CASE(RR)                                                                  \
    case RR:                                                                         \
        return forward ? fused3d_launch_r<T, RR, true>(pl, first, s, nparts)         \
                       : fused3d_launch_r<T, RR, false>(pl, first, s, nparts);
    switch (r) {
        F3_CASE(1)
        F3_CASE(2)
        F3_CASE(3)
        F3_CASE(4)
        F3_CASE(5)
        F3_CASE(6)
    }
#undef F3_CASE
    return nsol_fail(pl->ctx, NSOL_EINVAL, "fused3d: radius %d not instantiated", r);
}
