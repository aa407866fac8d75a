// pd_bulk_kernel.cuh -- 3-D fused primal-dual iteration with TMA bulk-async staged tiles.
//
// Same arithmetic, thread mapping and z-marching as pd_iter_kernel (pd_kernels.cu), but the
// global loads do not go through registers: one elected lane of every warp issues 1-D bulk copies
// (cp.async.bulk.shared.global, SASS UBLKCP -- the TMA unit) of the warp's contiguous row
// segments of xbar/p/x/b for plane z+S-1 into a warp-private ring of S shared-memory stages,
// completion is signalled on an mbarrier (expect_tx / complete_tx), and all lanes read plane z
// from shared memory.  Compared with the register-pipelined kernel this removes the per-thread
// 64-bit address arithmetic and predicates of 6-10 loads per plane and two in-flight register
// sets (float64: 164 -> ~90 registers), which is what bounds the float64 kernel.
// Row segments carry a 16-byte extension on each side (inside the domain) so the x-halos
// (xbar[x0-1], xbar[x0+W], p_x[x0-1]) arrive with the row; the y-halo rows of the CTA's edge
// warps are extra bulk copies.  The xbar / p'_y exchange between the warps of a CTA is unchanged
// (shared memory, one __syncthreads per plane).
// The warp's row index is broadcast from lane 0 so the compiler keeps the whole issue path (addresses,
// byte counts, predicates) on the uniform datapath; template parameters: LINK = in-kernel z-slab halo
// exchange over peer memory (pd_kernels.cu), UNIT = unit grid spacing (the w * a products are exact and dropped).
#pragma once

#ifndef NSOL_PD_STAGES
#define NSOL_PD_STAGES 3
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp; the compiler then issues the uniform-datapath TMA instructions
// under a plain predicate instead of an elect loop per instruction
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA, 1-D): 16-byte aligned src/dst, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared-memory plan of one CTA (TY warps, tile width W = 32*VEC, S stages)
template <typename T, int VEC>
struct PdBulkLayout {
    static constexpr int W = 32 * VEC;
    static constexpr int WX = W + 2 * VEC;                      // row with 16-byte extensions
    static constexpr int ROW_STAGE = 2 * WX + 4 * W;            // xbn, px (extended) + py, pz, x, b
    static constexpr int HALO_STAGE = 3 * W;                    // hup, hdn, pydn (CTA edge rows)
    __host__ __device__ static size_t bytes(int ty, int stages) {
        return 256 + sizeof(T) * ((size_t)(2 * (ty + 2) + 2 * (ty + 1)) * W + (size_t)stages * (ty * ROW_STAGE + HALO_STAGE));
    }
};

template <typename T, int VEC, int REG, int DATA, bool LINK, bool UNIT>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 2 : 1) pd_iter_bulk_kernel(const PdArgs<T> a) {
    using V = Vec<T, VEC>;
    using L = PdBulkLayout<T, VEC>;
    constexpr int S = NSOL_PD_STAGES;
    constexpr int W = L::W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_launch_dependents();

    const int lane = threadIdx.x;
    const int TY = (int)blockDim.y;
    const int ty = __shfl_sync(0xffffffffu, (int)threadIdx.y, 0);   // broadcast: the compiler then knows it is warp-uniform
    const int x0t = (int)blockIdx.x * W;                        // first voxel of the tile row
    const int x0 = x0t + lane * VEC;
    const int y = (int)blockIdx.y * TY + ty;
    int chunk = a.chunk_first + (int)(blockIdx.z % (unsigned)a.nsel) * a.chunk_stride;
    if (LINK && a.front_chunks == 1) chunk = chunk == 0 ? 0 : (chunk == 1 ? a.nchunks - 1 : chunk - 1);
    if (LINK && a.front_chunks == 2) chunk = chunk + 2 < a.nchunks ? chunk + 1 : (chunk + 2 == a.nchunks ? 0 : a.nchunks - 1);
    const int bz = (int)(blockIdx.z / (unsigned)a.nsel);
    const int z0 = chunk * a.zc;
    const int z1 = min(a.nz, z0 + a.zc);
    const bool row_in = (y < a.ny) && (x0t < a.nx);             // warp-uniform
    const bool active = row_in && (x0 < a.nx);

    const double *srow = a.sched + ((long long)a.it * a.batch + bz) * 8;
    const T sigma = (T)srow[0], tau = (T)srow[1], tl = (T)srow[2], theta = (T)srow[3];
    const ConstDiv<T> div_g((T)srow[4]), div_f((T)srow[5]);
    const T wx = a.wx, wy = a.wy, wz = a.wz;

    const long long sz = (long long)a.nx * a.ny;
    const long long trow = (long long)y * a.nx + x0t;           // tile-row offset inside a plane
    const long long hrow_t = (long long)bz * sz + trow;         // ... inside a halo plane array
    long long offt = (long long)bz * a.n + (long long)z0 * sz + trow;          // tile row, plane z
    long long bofft = (long long)bz * a.b_stride + (long long)z0 * sz + trow;
    const int lcol = lane * VEC;

    // ---- shared memory carve-up -------------------------------------------------------
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);    // [TY][S]
    T *s_xb = reinterpret_cast<T *>(smem_raw + 256);            // [2][TY+2][W]
    T *s_py = s_xb + 2 * (TY + 2) * W;                          // [2][TY+1][W]
    T *s_rows = s_py + 2 * (TY + 1) * W;                        // [S][TY][ROW_STAGE]
    T *s_halo = s_rows + S * TY * L::ROW_STAGE;                 // [S][HALO_STAGE]

    // tile geometry (warp-uniform)
    const int w_in = min(W, a.nx - x0t);                        // voxels of the row inside the domain
    const int ext_l = (x0t > 0) ? VEC : 0;
    const int ext_r = (x0t + W < a.nx) ? VEC : 0;
    const bool up_warp = ty == TY - 1, dn_warp = ty == 0;
    const bool need_up = up_warp && row_in && (y + 1 < a.ny);
    const bool need_dn = dn_warp && row_in && (y > 0);
    const bool need_r = active && lane == 31 && ext_r;          // lane 31 is inside the domain iff ext_r
    const bool need_l = active && lane == 0 && ext_l;
    const bool top_zero = (a.halo_xbar_above == nullptr);       // plane nz is the zero boundary

    if (threadIdx.x == 0 && threadIdx.y == 0) {
        for (int i = 0; i < TY * S; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pd_chain_begin(a, chunk);       // nothing above read solver state: the previous iteration may still be running

    // lane 0: bulk copies of everything plane zq needs into ring slot `slot`
    auto issue = [&](int zq, int slot, long long o, long long bo) {
        uint64_t *bar = &bars[ty * S + slot];
        T *st = s_rows + (slot * TY + ty) * L::ROW_STAGE;
        T *sh = s_halo + slot * L::HALO_STAGE;
        const unsigned row_b = (unsigned)(w_in * sizeof(T));
        const bool xbn_plane = zq + 1 < a.nz;                   // xbar plane zq+1 exists locally
        const bool xbn_halo = !LINK && !xbn_plane && !top_zero;  // ... or comes from the rank above (NCCL-filled
                                                                // buffer; in link mode it is read directly, below)
        const bool nx_plane = zq + 1 < z1;                      // CTA processes plane zq+1: needs its halos
        unsigned total = 5 * row_b + (unsigned)(ext_l * sizeof(T));          // px(+left ext), py, pz, x, b
        if (xbn_plane) total += row_b + (unsigned)((ext_l + ext_r) * sizeof(T));
        if (xbn_halo) total += row_b;
        if (need_up && nx_plane) total += row_b;
        if (need_dn && nx_plane) total += row_b;
        if (need_dn) total += row_b;
        mbar_expect_tx(bar, total);
        if (xbn_plane) bulk_g2s(st + VEC - ext_l, a.xbar_in + o + sz - ext_l, row_b + (unsigned)((ext_l + ext_r) * sizeof(T)), bar);
        if (xbn_halo) bulk_g2s(st + VEC, a.halo_xbar_above + hrow_t, row_b, bar);
        bulk_g2s(st + L::WX + VEC - ext_l, a.px_in + o - ext_l, row_b + (unsigned)(ext_l * sizeof(T)), bar);
        bulk_g2s(st + 2 * L::WX, a.py_in + o, row_b, bar);
        bulk_g2s(st + 2 * L::WX + W, a.pz_in + o, row_b, bar);
        bulk_g2s(st + 2 * L::WX + 2 * W, a.x + o, row_b, bar);
        bulk_g2s(st + 2 * L::WX + 3 * W, a.b + bo, row_b, bar);
        if (need_up && nx_plane) bulk_g2s(sh, a.xbar_in + o + sz + a.nx, row_b, bar);
        if (need_dn && nx_plane) bulk_g2s(sh + W, a.xbar_in + o + sz - a.nx, row_b, bar);
        if (need_dn) bulk_g2s(sh + 2 * W, a.py_in + o - a.nx, row_b, bar);
    };

    // ---- prologue ------------------------------------------------------------------------
    if (row_in && elect_one()) {
        for (int k = 0; k < S - 1; ++k)
            if (z0 + k < z1) issue(z0 + k, k, offt + (long long)k * sz, bofft + (long long)k * sz);
    }
    // xbar of plane z0 (and its halos) straight from global memory, once per chunk
    const long long off0 = offt + lcol;
    V xb_c = active ? vec_load<T, VEC>(a.xbar_in + off0) : vec_zero<T, VEC>();
    T xr_c = need_r ? a.xbar_in[off0 + VEC] : T(0);
    T xl_c = need_l ? a.xbar_in[off0 - 1] : T(0);
    V pz_prev = vec_zero<T, VEC>();
    if (LINK) pd_link_begin(a, z0, z1);
    if (active && (z0 > 0 || a.halo_pz_below)) {
        V xb_m = z0 > 0 ? vec_load<T, VEC>(a.xbar_in + off0 - sz) : vec_load_cg<T, VEC>(a.halo_xbar_below + hrow_t + lcol);
        V pz_m = z0 > 0 ? vec_load<T, VEC>(a.pz_in + off0 - sz) : vec_load_cg<T, VEC>(a.halo_pz_below + hrow_t + lcol);
#pragma unroll
        for (int v = 0; v < VEC; ++v) pz_prev.v[v] = dual_update<T, REG, UNIT>(pz_m.v[v], xb_c.v[v], xb_m.v[v], wz, sigma, div_g);
    }
    {
        T *buf = s_xb + (z0 & 1) * (TY + 2) * W;
        vec_store<T, VEC>(buf + (ty + 1) * W + lcol, xb_c);
        if (up_warp) {
            V h = (need_up && active) ? vec_load<T, VEC>(a.xbar_in + off0 + a.nx) : vec_zero<T, VEC>();
            vec_store<T, VEC>(buf + (TY + 1) * W + lcol, h);
        }
        if (dn_warp) {
            V h = (need_dn && active) ? vec_load<T, VEC>(a.xbar_in + off0 - a.nx) : vec_zero<T, VEC>();
            vec_store<T, VEC>(buf + lcol, h);
        }
        __syncthreads();
    }

    // ---- march through the chunk -------------------------------------------------------------
    int slot = 0;
    unsigned parity = 0;
    for (int z = z0; z < z1; ++z) {
        const bool more = (z + 1 < z1);
        // refill the slot consumed one plane ago (all its readers passed the last __syncthreads)
        if (row_in && (z + S - 1 < z1) && elect_one()) {
            int ps = slot + S - 1;
            if (ps >= S) ps -= S;
            issue(z + S - 1, ps, offt + (long long)(S - 1) * sz, bofft + (long long)(S - 1) * sz);
        }
        const T *st = s_rows + (slot * TY + ty) * L::ROW_STAGE;
        const T *sh = s_halo + slot * L::HALO_STAGE;
        V pxv = vec_zero<T, VEC>(), pyv = pxv, pzv = pxv, xv = pxv, bv = pxv, xbn = pxv, hup = pxv, hdn = pxv, pydn = pxv;
        T pxl = T(0), xr_n = T(0), xl_n = T(0);
        if (row_in) {
            mbar_wait(&bars[ty * S + slot], parity);
            if (active) {
                if (z + 1 < a.nz || (!LINK && !top_zero)) xbn = vec_load<T, VEC>(st + VEC + lcol);
                pxv = vec_load<T, VEC>(st + L::WX + VEC + lcol);
                pyv = vec_load<T, VEC>(st + 2 * L::WX + lcol);
                pzv = vec_load<T, VEC>(st + 2 * L::WX + W + lcol);
                xv = vec_load<T, VEC>(st + 2 * L::WX + 2 * W + lcol);
                bv = vec_load<T, VEC>(st + 2 * L::WX + 3 * W + lcol);
                if (need_l) pxl = st[L::WX + VEC - 1];
                if (more) {
                    if (need_r) xr_n = st[VEC + W];
                    if (need_l) xl_n = st[VEC - 1];
                    if (need_up) hup = vec_load<T, VEC>(sh + lcol);
                    if (need_dn) hdn = vec_load<T, VEC>(sh + W + lcol);
                }
                if (need_dn) pydn = vec_load<T, VEC>(sh + 2 * W + lcol);
            }
        }

        if (LINK && z + 1 == a.nz && !top_zero) {
            // xbar plane of the rank above: stored by that rank's kernel into this rank's link block
            if (active) xbn = vec_load_cg<T, VEC>(a.halo_xbar_above + hrow_t + lcol);
        }

        // neighbours of the current plane
        T *xbuf_c = s_xb + (z & 1) * (TY + 2) * W;
        T *xbuf_n = s_xb + ((z + 1) & 1) * (TY + 2) * W;
        T *pbuf = s_py + (z & 1) * (TY + 1) * W;
        V xup = vec_load<T, VEC>(xbuf_c + (ty + 2) * W + lcol);
        V xdn = vec_zero<T, VEC>();
        if (dn_warp) xdn = vec_load<T, VEC>(xbuf_c + lcol);
        T x_right = shfl_down_t(xb_c.v[0], 1);
        if (lane == 31) x_right = xr_c;

        // dual update
        V pnx, pny, pnz;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            T hi = (v + 1 < VEC) ? xb_c.v[(v + 1) % VEC] : x_right;
            pnx.v[v] = dual_update<T, REG, UNIT>(pxv.v[v], hi, xb_c.v[v], wx, sigma, div_g);
            pny.v[v] = dual_update<T, REG, UNIT>(pyv.v[v], xup.v[v], xb_c.v[v], wy, sigma, div_g);
            pnz.v[v] = dual_update<T, REG, UNIT>(pzv.v[v], xbn.v[v], xb_c.v[v], wz, sigma, div_g);
        }
        const long long off = offt + lcol;
        if (active) {
            vec_store<T, VEC>(a.px_out + off, pnx);
            vec_store<T, VEC>(a.py_out + off, pny);
            vec_store<T, VEC>(a.pz_out + off, pnz);
        }
        T pnx_left = shfl_up_t(pnx.v[VEC - 1], 1);
        if (lane == 0) pnx_left = need_l ? dual_update<T, REG, UNIT>(pxl, xb_c.v[0], xl_c, wx, sigma, div_g) : T(0);

        // publish p'_y of this row (and of the halo row below the tile) and the next xbar plane
        vec_store<T, VEC>(pbuf + (ty + 1) * W + lcol, pny);
        if (dn_warp) {
            V h = vec_zero<T, VEC>();
            if (need_dn && active) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) h.v[v] = dual_update<T, REG, UNIT>(pydn.v[v], xb_c.v[v], xdn.v[v], wy, sigma, div_g);
            }
            vec_store<T, VEC>(pbuf + lcol, h);
        }
        if (more) {
            vec_store<T, VEC>(xbuf_n + (ty + 1) * W + lcol, xbn);
            if (up_warp) vec_store<T, VEC>(xbuf_n + (TY + 1) * W + lcol, hup);
            if (dn_warp) vec_store<T, VEC>(xbuf_n + lcol, hdn);
        }
        __syncthreads();
        V pny_dn = vec_load<T, VEC>(pbuf + ty * W + lcol);

        // primal update + over-relaxation
        V xnew, xbnew;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            T lo = (v == 0) ? pnx_left : pnx.v[(v + VEC - 1) % VEC];
            T div = wdiff<T, UNIT>(wx, lo, pnx.v[v]);
            div = div + wdiff<T, UNIT>(wy, pny_dn.v[v], pny.v[v]);
            div = div + wdiff<T, UNIT>(wz, pz_prev.v[v], pnz.v[v]);
            primal_update<T, DATA>(xv.v[v], bv.v[v], div, tau, tl, theta, div_f, xnew.v[v], xbnew.v[v]);
        }
        if (active) {
            vec_store<T, VEC>(a.x + off, xnew);
            vec_store<T, VEC>(a.xbar_out + off, xbnew);
        }

        xb_c = xbn;
        pz_prev = pnz;
        xr_c = xr_n;
        xl_c = xl_n;
        offt += sz;
        bofft += sz;
        if (++slot == S) {
            slot = 0;
            parity ^= 1u;
        }
    }
    if (LINK) pd_link_finish<T, VEC>(a, z0, z1, hrow_t + lcol, active);
    pd_chain_end(a, chunk);
}
