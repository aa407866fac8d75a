/* nsol_b200.h -- C ABI of libnsol_b200.so: the B200 (sm_100a) backend for the
 * iterative proximal-solver hot path of gift-surg/NSoL.
 *
 * The reference (NSoL v0.1.14) is pure Python and has no FFI layer of its own;
 * the drop-in boundary is its Python object API.  Every entry
 * point below names the reference interface (file:line under the reference
 * root) whose body it replaces; the Python classes in nsol_b200/ keep the
 * reference's names and signatures and bind these symbols through ctypes
 * (nsol_b200/_lib.py).  INTEGRATION.md shows the stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes only, no C++/torch types.
 *   - every function returns NSOL_OK (0) or a negative nsol_status; the message
 *     is available from nsol_last_error().
 *   - volumes are C-order, numpy axis order: shape[0..dim-1] = (z, y, x) for
 *     dim 3, (y, x) for dim 2, (x) for dim 1.  spacing[k] divides derivative
 *     k, where k=0 acts on the LAST axis, k=1 on axis -2, k=2 on axis 0
 *     (nsol/kernels.py:160-190, 240-286).
 *   - the dual variable is SoA: dim blocks of N elements each,
 *     [D0 block | D1 block | D2 block]  (nsol/linear_operators.py:132,140).
 *   - "dev" pointers are device memory on the context's GPU, "host" pointers
 *     are ordinary (preferably pinned) host memory holding float64, the
 *     reference's only element type (nsol/solver.py:37).
 *   - stream is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Calls are asynchronous w.r.t. the host unless they return data to the host.
 *   - no internal threads; one context per host thread / GPU.
 *   - there is no CPU fallback: without a CUDA device nsol_create() fails.
 */
#ifndef NSOL_B200_H
#define NSOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSOL_B200_VERSION 100 /* 0.1.0 */

typedef struct nsol_ctx nsol_ctx;
typedef struct nsol_pd_plan nsol_pd_plan;
typedef struct nsol_lsmr_plan nsol_lsmr_plan;
typedef void *nsol_stream; /* cudaStream_t */

typedef enum {
    NSOL_OK = 0,
    NSOL_EINVAL = -1,  /* bad argument (ValueError on the Python side) */
    NSOL_ECUDA = -2,   /* CUDA runtime error */
    NSOL_ENOMEM = -3,  /* device / host allocation failed */
    NSOL_ESTATE = -4,  /* call sequence error (plan not reset, ...) */
    NSOL_ENCCL = -5    /* communicator error */
} nsol_status;

typedef enum { NSOL_F64 = 0, NSOL_F32 = 1 } nsol_dtype;
/* regulariser of the dual prox: nsol/proximal_operators.py:139-140 (TV), :157-159 (Huber);
 * TK1 = BASELINE config 5's callable  q/(1+sigma). */
typedef enum { NSOL_REG_TV = 0, NSOL_REG_HUBER = 1, NSOL_REG_TK1 = 2 } nsol_reg;
/* data prox: nsol/proximal_operators.py:96-98 (L1), :118-120 (L2) */
typedef enum { NSOL_DATA_L1 = 0, NSOL_DATA_L2 = 1 } nsol_data;
/* step-size schedules: nsol/primal_dual_solver.py:278-306 (ALG2), :374-403 (AHMOD), :321-358 (ALG3) */
typedef enum { NSOL_ALG2 = 0, NSOL_ALG2_AHMOD = 1, NSOL_ALG3 = 2 } nsol_alg;
/* second operator of the stacked least-squares system */
typedef enum { NSOL_B_GRAD = 0, NSOL_B_IDENTITY = 1, NSOL_B_NONE = 2 } nsol_bop;
typedef enum { NSOL_A_BLUR = 0, NSOL_A_IDENTITY = 1 } nsol_aop;

/* Regular grid + element type + number of independent problems stacked
 * contiguously (batch stride = number of voxels). */
typedef struct {
    int32_t dim;        /* 1, 2 or 3 */
    int32_t dtype;      /* nsol_dtype: arithmetic / storage type on the device */
    int64_t shape[3];   /* numpy order, entries >= dim ignored */
    double spacing[3];  /* entries >= dim ignored */
    int32_t batch;      /* >= 1 */
    int32_t reserved;
} nsol_grid;

/* ---- library / context ------------------------------------------------- */
int nsol_version(void);
/* device < 0: use the current device.  Fails (NSOL_ECUDA) without a GPU. */
int nsol_create(int device, nsol_ctx **out);
void nsol_destroy(nsol_ctx *ctx);
/* ctx may be NULL: returns the calling thread's last creation error. */
const char *nsol_last_error(const nsol_ctx *ctx);
/* tuning knobs ("pd_zc", "pd_ty", "pd_variant", "pd_persist", "pd_pipe", "pd_pipe_depth", "pd_pipe_planes", "pd_tb", "pd_tb_k", "pd_tb_nr", "pd_pdl", "pd_chain", "pd_push", "lsmr_blocks", "lsmr_path", "lsmr_fuse2d", "lsmr_fuse3d", "lsmr_tile", "link_timeout_ms", "debug_guard"); value <= 0 restores the default */
int nsol_set_tuning(nsol_ctx *ctx, const char *key, int value);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t nsol_launch_count(const nsol_ctx *ctx);
int nsol_device_sm_count(const nsol_ctx *ctx);
/* Debug aid (the pool's compute-sanitizer is closed): after nsol_set_tuning(ctx, "debug_guard", 1) the arrays of every plan
 * created on this context sit between two 64 KiB guard bands filled with NaN bit patterns -- an out-of-bounds read drags NaN
 * into the result, an out-of-bounds write is counted here.  Synchronises the device.  violations_out: guard bytes overwritten
 * so far; arrays_out (may be NULL): guarded arrays currently alive. */
int nsol_debug_guard_check(nsol_ctx *ctx, int64_t *violations_out, int *arrays_out);

/* ---- memory / stream helpers (plumbing for hosts without a CUDA binding) - */
int nsol_device_alloc(nsol_ctx *ctx, size_t bytes, void **dev);
int nsol_device_free(nsol_ctx *ctx, void *dev);
int nsol_host_alloc(nsol_ctx *ctx, size_t bytes, void **host); /* pinned */
int nsol_host_free(nsol_ctx *ctx, void *host);
int nsol_memcpy_h2d(nsol_ctx *ctx, void *dev, const void *host, size_t bytes, nsol_stream s);
int nsol_memcpy_d2h(nsol_ctx *ctx, void *host, const void *dev, size_t bytes, nsol_stream s);
int nsol_memcpy_d2d(nsol_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes, nsol_stream s);
int nsol_memset_dev(nsol_ctx *ctx, void *dev, int value, size_t bytes, nsol_stream s);
int nsol_stream_sync(nsol_ctx *ctx, nsol_stream s);
/* out[i] = (T_out)(in[i] * mul)  resp.  (T_out)(in[i] / div); dtypes are nsol_dtype */
int nsol_scale_convert(nsol_ctx *ctx, int64_t n, int dtype_in, const void *in_dev, int dtype_out,
                       void *out_dev, double factor, int divide, nsol_stream s);

/* ---- linear operators on device arrays ----------------------------------
 * nsol_grad      replaces LinearOperators.get_gradient_operators()[0]
 *                (nsol/linear_operators.py:121-144; D_k from :98-106,:193-201,:219-247)
 * nsol_grad_adj  replaces get_gradient_operators()[1] (:158-169)
 * nsol_diff / nsol_diff_adj  a single D_k / D_k^T (get_dx/dy/dz_operators)
 * nsol_blur_sep  replaces get_gaussian_blurring_operators() for a separable
 *                (diagonal covariance) mask, periodic boundary (:82-86, :60-68);
 *                taps[a]/radius[a] belong to numpy axis a; tmp_dev holds N*batch
 *                elements of scratch (may equal NULL for dim 1).
 * nsol_conv_wrap replaces get_convolution_and_adjoint_convolution_operators()
 *                for an arbitrary dense odd-sized mask (:60-68).
 * x: N*batch elements; p: dim*N per problem, problems contiguous. */
int nsol_grad(nsol_ctx *ctx, const nsol_grid *g, const void *x_dev, void *p_dev, nsol_stream s);
int nsol_grad_adj(nsol_ctx *ctx, const nsol_grid *g, const void *p_dev, void *x_dev, nsol_stream s);
int nsol_diff(nsol_ctx *ctx, const nsol_grid *g, int component, int adjoint, const void *in_dev,
              void *out_dev, nsol_stream s);
int nsol_blur_sep(nsol_ctx *ctx, const nsol_grid *g, const double *const taps_host[3],
                  const int32_t radius[3], const void *x_dev, void *y_dev, void *tmp_dev,
                  nsol_stream s);
int nsol_conv_wrap(nsol_ctx *ctx, const nsol_grid *g, const double *kernel_host,
                   const int64_t kshape[3], const void *x_dev, void *y_dev, nsol_stream s);

/* ---- stand-alone proximal maps on device arrays --------------------------
 * Replace the static methods of ProximalOperators when they are called on
 * arrays directly (nsol/proximal_operators.py:96-98, 118-120, 139-140, 157-159).
 *   TV_CONJ    out = x / max(1,|x|)
 *   HUBER_CONJ out = y / max(1,|y|),  y = x / (1 + p0*p1)      (p0 = sigma, p1 = gamma)
 *   TK1_CONJ   out = x / (1 + p0)
 *   ELL1       out = b + max(|x-b| - p0, 0) sign(x-b),  b = x0 / p1   (p0 = tau, p1 = x_scale)
 *   ELL2       out = (x + p0 b) / (1 + p0)
 * x0_dev is only read by ELL1 / ELL2.  n elements of grid dtype `dtype`. */
typedef enum { NSOL_PROX_TV_CONJ = 0, NSOL_PROX_HUBER_CONJ = 1, NSOL_PROX_TK1_CONJ = 2,
               NSOL_PROX_ELL1 = 3, NSOL_PROX_ELL2 = 4 } nsol_prox_kind;
int nsol_prox_apply(nsol_ctx *ctx, int kind, int dtype, int64_t n, const void *x_dev, const void *x0_dev,
                    double p0, double p1, void *out_dev, nsol_stream s);

/* ---- on-device measures ----------------------------------
 * One reduction pass instead of copying an iterate to the host.  y = x * scale (the solver's
 * x_scale), r = reference image (float64, n elements):
 *   out[0..7] = sum y, sum y^2, sum r, sum r^2, sum y r, sum (y-r)^2, sum |y-r|, max r
 * from which SSD / MSE / RMSE / MAE / PSNR / NCC follow (nsol/similarity_measures.py:26-120).
 * nsol_prior_stats: out[0] = sum_i sqrt(sum_k (D_k y)_i^2)   (total variation, nsol/prior_measures.py:27-37)
 *                   out[1] = 1/2 sum_i sum_k (D_k y)_i^2     (first-order Tikhonov, :22-24)
 *                   out[2] = sum_i huber(|grad y|_i^2, gamma) / (2 gamma)   (:40-52)
 *                   out[3] = 1/2 sum_i y_i^2                 (zeroth-order Tikhonov, :19-20)
 * Results are written to host memory; the call synchronises the stream. */
int nsol_similarity_stats(nsol_ctx *ctx, int dtype_x, int64_t n, const void *x_dev, double scale,
                          const double *xref_dev, double *out_host8, nsol_stream s);
int nsol_prior_stats(nsol_ctx *ctx, const nsol_grid *g, const void *x_dev, double scale, double huber_gamma,
                     double *out_host4, nsol_stream s);

/* ---- fused primal-dual (Chambolle-Pock) ----------------------------------
 * Replaces the body of PrimalDualSolver._run (nsol/primal_dual_solver.py:215-263)
 * for B = grad, B_conj = grad_adj, prox_g_conj in {tv, huber, tk1} and
 * prox_f in {ell1, ell2 denoising}.  One launch per iteration:
 *     p    <- prox_g*(p + sigma grad(xbar))
 *     x+   <- prox_f(x - tau grad_adj(p), tau/alpha)
 *     xbar <- x+ + theta (x+ - x)
 * Step sizes follow :278-403 and are evaluated on the host in float64.
 * alpha has one entry per batch member (a parameter sweep shares one
 * observation: b_batched = 0; an image batch has b_batched = 1). */
typedef struct {
    nsol_grid grid;
    int32_t reg;        /* nsol_reg */
    int32_t data;       /* nsol_data */
    int32_t alg;        /* nsol_alg */
    int32_t b_batched;  /* 0: one observation shared by the batch, 1: batch observations */
    double huber_gamma; /* nsol/proximal_operators.py:157 default 0.05 */
    double L2;          /* squared operator norm handed to the solver */
    double x_scale;     /* get_x() multiplies by it (nsol/solver.py:117-118) */
    double x0_scale;    /* x = xbar = x0 / x0_scale (nsol/solver.py:35-41); 1 if x0 is already scaled */
    double b_scale;     /* b' = b / b_scale: the x_scale argument of prox_ell*_denoising (proximal_operators.py:97,119) */
    const double *alpha; /* [grid.batch] */
} nsol_pd_desc;

int nsol_pd_plan_create(nsol_ctx *ctx, const nsol_pd_desc *desc, nsol_pd_plan **out);
void nsol_pd_plan_destroy(nsol_pd_plan *plan);
/* change the solver parameters (reg, data, alg, gamma, L2, scales, alpha) of an existing plan
 * without reallocating; grid, dtype and batch must be unchanged.  The plan must be reset again. */
int nsol_pd_plan_update(nsol_pd_plan *plan, const nsol_pd_desc *desc);
/* bytes of device memory owned by the plan */
size_t nsol_pd_plan_bytes(const nsol_pd_plan *plan);
/* (re)start: x = xbar = x0/x0_scale, b' = b/b_scale, p = 0, iteration counter = 0.
 * x0 may equal b (or be NULL = b).  *_host take float64 host arrays,
 * *_dev take device arrays of the plan's dtype. */
int nsol_pd_plan_reset_host(nsol_pd_plan *plan, const double *b_host, const double *x0_host, nsol_stream s);
int nsol_pd_plan_reset_dev(nsol_pd_plan *plan, const void *b_dev, const void *x0_dev, nsol_stream s);
/* advance by n iterations (asynchronous) */
int nsol_pd_plan_iterate(nsol_pd_plan *plan, int n, nsol_stream s);
int nsol_pd_plan_iterations_done(const nsol_pd_plan *plan);
/* current primal iterate in solver units (x / x_scale), plan dtype, N*batch elements */
int nsol_pd_plan_x_dev(nsol_pd_plan *plan, const void **x_dev);
/* get_x(): x * x_scale as float64 into host memory (synchronises the stream) */
int nsol_pd_plan_get_x_host(nsol_pd_plan *plan, double *x_host, nsol_stream s);
/* device-resident variant of the same: x * x_scale into a device array of dtype_out */
int nsol_pd_plan_get_x_dev(nsol_pd_plan *plan, int dtype_out, void *out_dev, nsol_stream s);
/* One whole solve of an existing plan from / to host float64 buffers: reset with (b_host, x0_host; x0_host NULL = b_host),
 * `iterations` iterations, x_host = x * x_scale.  Replaces PrimalDualSolver.run() + get_x() without an Observer
 * (nsol/primal_dual_solver.py:215-263, nsol/solver.py:117-118).  For a large single volume with page-locked host buffers the
 * upload / download are cut into groups of z-planes that overlap a wavefront of iterations (tuning knobs "pd_pipe",
 * "pd_pipe_depth", "pd_pipe_planes"); otherwise it is reset_host + iterate + get_x_host.  Same kernels, same bits.  Synchronous. */
int nsol_pd_plan_solve_host(nsol_pd_plan *plan, const double *b_host, const double *x0_host, int iterations, double *x_host,
                            nsol_stream s);
/* linked z-slabs (nsol_pd_plan_link_*): direction of the transfer groups of nsol_pd_plan_solve_host, +1 bottom-up (default),
 * -1 top-down; neighbouring slabs must alternate (rank parity) so that the wavefront continues through the slab boundaries */
int nsol_pd_plan_set_pipe_direction(nsol_pd_plan *plan, int direction);
/* transfer groups / wavefront depth used by the last nsol_pd_plan_solve_host (0 groups: the plain sequence) */
int nsol_pd_plan_solve_info(const nsol_pd_plan *plan, int *groups_out, int *depth_out);
/* one call = PrimalDualSolver.run() + get_x() with host float64 buffers
 * (H2D, iterations, D2H).  iterates_host, if not NULL, receives the
 * (iterations+1) arrays the Observer would see (:218-219, :260-261). */
int nsol_pd_run_host(nsol_ctx *ctx, const nsol_pd_desc *desc, int iterations, const double *b_host,
                     const double *x0_host, double *x_host, double *iterates_host, nsol_stream s);
/* z-slab decomposition (one rank per GPU): the plan's volume is the local slab
 * [z_lo, z_hi) of a taller volume.  halo_* are device arrays of one plane
 * (ny*nx*batch elements) owned by the caller and refreshed by the caller's
 * exchange (NCCL send/recv) before every iteration:
 *   xbar_above : xbar plane z_hi of the upper neighbour  (NULL at the global top: Dirichlet 0)
 *   xbar_below / p_below : xbar and p_z plane z_lo-1 of the lower neighbour (NULL at the bottom)
 * nsol_pd_plan_boundary_planes exposes the plan's own boundary planes to send. */
int nsol_pd_plan_set_halo(nsol_pd_plan *plan, const void *xbar_above, const void *xbar_below,
                          const void *p_below);
int nsol_pd_plan_boundary_planes(nsol_pd_plan *plan, const void **xbar_first, const void **xbar_last,
                                 const void **pz_last);
/* Split iteration, to overlap the halo exchange with compute: part 1 launches only the first and
 * last z-chunk (afterwards nsol_pd_plan_boundary_planes_next returns the planes of the NEW state to
 * send), part 2 the interior chunks and advances the plan.  Needs nsol_pd_plan_chunks() >= 3. */
int nsol_pd_plan_iterate_part(nsol_pd_plan *plan, int part, nsol_stream s);
int nsol_pd_plan_boundary_planes_next(nsol_pd_plan *plan, const void **xbar_first, const void **xbar_last,
                                      const void **pz_last);
int nsol_pd_plan_chunks(nsol_pd_plan *plan);
/* In-kernel halo exchange over peer memory (NVLink / NVSwitch) -- the fused form of the z-slab
 * exchange: no NCCL call and no host work per iteration.  Every rank owns a "link block" (flags +
 * double-buffered receive slots for the three halo planes).  The boundary CTAs of an iteration
 * store the first xbar plane / the last xbar and p_z planes of the state they produce straight
 * into the neighbours' blocks and raise a flag there (release, system scope); the next iteration
 * of the neighbour spins on that flag (acquire) before it reads the halo.  With a link connected,
 * nsol_pd_plan_iterate(plan, n) just queues n launches; all ranks must call reset / iterate in
 * lockstep (same sequence, same n).  The first iterate after a reset also publishes the boundary
 * planes of the start state (n = 0 does only that -- needed when one stream drives several linked
 * plans).  nsol_pd_plan_set_halo is not used in this mode.
 *   link_create      allocates this plan's block and returns its device pointer / size
 *   link_ipc_handle  64-byte cudaIpcMemHandle_t of the block, to be sent to the neighbours
 *   link_open        maps the neighbours' blocks from their IPC handles (NULL = no neighbour:
 *                    global bottom / top) and switches the plan to link mode
 *   link_connect     same with raw device pointers valid in this process (one process driving
 *                    several plans / GPUs; used by the single-GPU emulation tests)
 *   link_status      synchronises and returns NSOL_ENCCL if a flag wait timed out */
int nsol_pd_plan_link_create(nsol_pd_plan *plan, void **block_dev, size_t *block_bytes);
int nsol_pd_plan_link_ipc_handle(nsol_pd_plan *plan, void *handle64);
int nsol_pd_plan_link_open(nsol_pd_plan *plan, const void *handle_below64, const void *handle_above64);
int nsol_pd_plan_link_connect(nsol_pd_plan *plan, void *block_below, void *block_above);
int nsol_pd_plan_link_status(nsol_pd_plan *plan, nsol_stream s);

/* ---- stacked least squares: LSMR on [A; sqrt(alpha) B] ---------------------
 * Replaces TikhonovLinearSolver._run, lsmr/linear branch
 * (nsol/tikhonov_linear_solver.py:120-158, 226-274) including
 * scipy.sparse.linalg.lsmr(A, b, maxiter, atol=0, btol=0) (cold start) and the
 * final clip to [lo, hi]; and, on top of it, ADMMLinearSolver._run
 * (nsol/admm_linear_solver.py:165-253). */
typedef struct {
    nsol_grid grid;      /* batch must be 1 */
    int32_t a_op;        /* nsol_aop */
    int32_t b_op;        /* nsol_bop */
    const double *taps[3]; /* separable blur taps per numpy axis (a_op == BLUR) */
    int32_t radius[3];
    int32_t reserved;
} nsol_lsq_desc;

int nsol_lsmr_plan_create(nsol_ctx *ctx, const nsol_lsq_desc *desc, nsol_lsmr_plan **out);
void nsol_lsmr_plan_destroy(nsol_lsmr_plan *plan);
size_t nsol_lsmr_plan_bytes(const nsol_lsmr_plan *plan);
/* x_dev <- clip(lsmr([A; sqrt(alpha) B], [b; sqrt(alpha) b_reg], maxiter), lo, hi).
 * b_dev: N elements; b_reg_dev: rows of B (dim*N for grad, N for identity; NULL = 0).
 * Device arrays of the plan's dtype, already in solver units.  itn_out/istop_out may be NULL. */
int nsol_lsmr_solve_dev(nsol_lsmr_plan *plan, double alpha, const void *b_dev, const void *b_reg_dev,
                        int maxiter, double lo, double hi, void *x_dev, int *itn_out, int *istop_out,
                        nsol_stream s);
/* TikhonovLinearSolver.run() + get_x() with host float64 buffers: the solve uses
 * b / in_scale and b_reg / in_scale (nsol/linear_solver.py:83, tikhonov_linear_solver.py:89;
 * pass 1 if the host already scaled them) and returns x * out_scale (nsol/solver.py:117-118). */
int nsol_tikhonov_run_host(nsol_lsmr_plan *plan, double alpha, double in_scale, double out_scale,
                           const double *b_host, const double *b_reg_host, int maxiter, double lo,
                           double hi, double *x_host, nsol_stream s);
/* ADMMLinearSolver.run() + get_x(): iterations outer steps, iter_max LSMR steps each.
 * iterates_host (may be NULL): (iterations+1)*N values as the Observer sees them. */
int nsol_admm_run_host(nsol_lsmr_plan *plan, double alpha, double rho, int iterations, int iter_max,
                       double in_scale, double out_scale, const double *b_host, const double *x0_host,
                       double *x_host, double *iterates_host, nsol_stream s);
/* ADMMLinearSolver(..., b_reg=...): the solver's own offset of the regulariser rows
 * (nsol/admm_linear_solver.py:100, used at :171 v0 = B x0 - b_reg, :208 t = B x + w - b_reg,
 * :222 Tikhonov b_reg = v - w + b_reg).  dim*N float64 host values, divided by in_scale (= x_scale);
 * kept by the plan for every following nsol_admm_run_*.  NULL restores the default b_reg = 0. */
int nsol_admm_set_b_reg_host(nsol_lsmr_plan *plan, const double *b_reg_host, double in_scale, nsol_stream s);
/* PrimalDualSolver.run() + get_x() for the deconvolution wiring
 * (nsol/deconvolution_solver_parameter_study_interface.py:255-280, 303-325):
 *   prox_g_conj in {tv, huber}, B = grad, and
 *   prox_f = prox_linear_least_squares(x, tau, A, A_adj, b, x0, iter_max, x_scale=prox_scale)
 * (nsol/proximal_operators.py:44-78): every iteration solves
 *   min_y 1/2 ||A y - b''||^2 + 1/(2 t) ||y - x'/prox_scale||^2,  b'' = b / prox_scale^2, t = tau/alpha
 * by LSMR (iter_max steps, cold start, clipped to [0, inf)) on [A; sqrt(1/t) I] and returns y * prox_scale.
 * The plan's b_op must be NSOL_B_IDENTITY.  pd.grid must equal the plan's grid (batch 1); pd.x_scale is
 * the solver's x_scale, x0_host is already divided by it (pd.x0_scale = 1); pd.b_scale is unused. */
int nsol_pd_deconv_run_host(nsol_lsmr_plan *plan, const nsol_pd_desc *pd, int iterations, int iter_max,
                            double prox_scale, const double *b_host /* raw observation of the prox */,
                            const double *x0_host, double *x_host, double *iterates_host, nsol_stream s);

/* ---- z-slab decomposition of the LSMR / ADMM path (one rank per GPU) -------------
 * The plan's volume is the local slab [z_lo, z_hi) of a taller volume (numpy axis 0).  The library does
 * no communication itself: the caller exchanges halo planes with the neighbouring ranks (NCCL send/recv)
 * and all-reduces one double per reduction; this API exposes the buffers and cuts a solve into phases at
 * every such point.  Blur halos are periodic (ring: the lower neighbour of rank 0 is the last rank,
 * nsol/linear_operators.py:60-68 mode="wrap"); the gradient keeps its zero boundary at the global ends
 * (has_below / has_above = 0 there, :98-106).
 *   nsol_lsmr_plan_slab     switch a plan to slab mode and allocate its halo buffers
 *   nsol_lsmr_slab_buffers  which = V | U0 | UZ | X: the local planes to send (first / last `planes`
 *                           planes of the array; NULL if that direction is not needed) and the receive
 *                           buffers (recv_lo <- lower neighbour's last planes, recv_hi <- upper
 *                           neighbour's first planes)
 *   nsol_lsmr_slab_arrays   b (N, solver units, written by the caller), x (N: start value in, result
 *                           out) and the one-double reduction buffer to all-reduce (sum) after every
 *                           phase that produces a sum of squares
 *   nsol_lsmr_slab_phase    one phase (asynchronous).  Sequence of a solve with weight alpha
 *                           (tikhonov_linear_solver.py:226-274 + scipy lsmr.py:239-479):
 *       RHS(p0 = sqrt(alpha), i0 = 1: use the plan's b_reg, 0: b_reg = 0) . allreduce .
 *       SCAL_INIT_BETA(p0 = sqrt(alpha), i0 = maxiter)
 *       [exchange U0, UZ] ADJ_FIRST . allreduce . SCAL_INIT_ALPHA
 *       maxiter x { [exchange V] FWD . allreduce . SCAL_BETA . [exchange U0, UZ] ADJ . allreduce .
 *                   SCAL_ALPHA . UPDATE . allreduce . SCAL_TESTS }
 *       CLIP(p0 = lo, p1 = hi)
 *     ADMM (admm_linear_solver.py:165-253): [exchange X] ADMM_INIT, then per outer iteration a solve
 *     with alpha = rho (b_reg = v - w is kept by the plan) followed by [exchange X] ADMM_SHRINK(p0 = alpha/rho).
 *   nsol_lsmr_plan_status   iteration count / istop of the last solve (synchronises) */
typedef enum { NSOL_SLAB_V = 0, NSOL_SLAB_U0 = 1, NSOL_SLAB_UZ = 2, NSOL_SLAB_X = 3 } nsol_slab_buffer;
typedef enum {
    NSOL_PH_RHS = 0, NSOL_PH_SCAL_INIT_BETA = 1, NSOL_PH_ADJ_FIRST = 2, NSOL_PH_SCAL_INIT_ALPHA = 3,
    NSOL_PH_FWD = 4, NSOL_PH_SCAL_BETA = 5, NSOL_PH_ADJ = 6, NSOL_PH_SCAL_ALPHA = 7, NSOL_PH_UPDATE = 8,
    NSOL_PH_SCAL_TESTS = 9, NSOL_PH_CLIP = 10, NSOL_PH_ADMM_INIT = 11, NSOL_PH_ADMM_SHRINK = 12
} nsol_slab_phase;
int nsol_lsmr_plan_slab(nsol_lsmr_plan *plan, int has_below, int has_above);
int nsol_lsmr_slab_buffers(nsol_lsmr_plan *plan, int which, const void **send_first, const void **send_last,
                           void **recv_lo, void **recv_hi, int *planes);
int nsol_lsmr_slab_arrays(nsol_lsmr_plan *plan, void **b_dev, void **x_dev, double **ss_dev);
int nsol_lsmr_slab_phase(nsol_lsmr_plan *plan, int phase, double p0, double p1, int i0, nsol_stream s);
int nsol_lsmr_plan_status(nsol_lsmr_plan *plan, int *itn_out, int *istop_out, nsol_stream s);

/* device-resident ADMM used by bench.py: arrays in solver units, plan dtype */
int nsol_admm_run_dev(nsol_lsmr_plan *plan, double alpha, double rho, int iterations, int iter_max,
                      const void *b_dev, const void *x0_dev, void *x_dev, nsol_stream s);
/* isotropic shrink of ADMMLinearSolver._prox_g (nsol/admm_linear_solver.py:239-253):
 * v = shrink(t, ell), w = t - v with t = B(x) + w_in  (w_in may be NULL = 0) */
int nsol_admm_shrink(nsol_ctx *ctx, const nsol_grid *g, const void *x_dev, const void *w_in_dev,
                     double ell, void *v_dev, void *w_dev, nsol_stream s);

#ifdef __cplusplus
}
#endif
#endif /* NSOL_B200_H */
