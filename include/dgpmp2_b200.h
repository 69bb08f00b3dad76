/*
 * dgpmp2_b200 -- C ABI of the B200-native dGPMP2 inner Gauss-Newton path.
 *
 * This is the drop-in boundary: everything the Python host side (the mirror of
 * the reference's diff_gpmp2.gpmp2 planner/factor API) needs from the GPU goes
 * through the entry points declared here.  Plain pointers and sizes only -- no
 * torch types.  All array arguments are DEVICE pointers to contiguous row-major
 * buffers owned by the caller unless an entry point says "host".  Every call
 * is asynchronous on the given CUDA stream (`stream` is a `cudaStream_t` passed
 * as `void*`; NULL = the legacy default stream), allocates nothing, keeps no
 * global state and is therefore thread-safe per stream.  Return value: 0 on
 * success, a negative DGPMP2_ERR_* code otherwise (nothing is launched then).
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; each entry
 * point below cites the reference function(s) (file:line under
 * /root/reference) whose work it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Suffix _f32 / _f64 = element type of the trajectory / SDF / weight / output
 * buffers ("I/O type").  Factor evaluation, right-hand sides, errors and residuals are IEEE
 * double inside every kernel; dgpmp2_gn_step_f32 factorises in fp32 and refines with double
 * residuals (falling back per problem to the all-double solver), everything else solves in double
 * ("numerics").
 */
#ifndef DGPMP2_B200_H
#define DGPMP2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGPMP2_ABI_VERSION 1

#define DGPMP2_OK 0
#define DGPMP2_ERR_ARG (-1)         /* null pointer / non-positive size / bad flag combination */
#define DGPMP2_ERR_UNSUPPORTED (-2) /* trajectory too long for on-chip band storage, dof not in {2,3} */
#define DGPMP2_ERR_CUDA (-3)        /* a CUDA runtime call failed; see dgpmp2_last_cuda_error() */

/* feature flags (reference plan_layer.py:32-36, :90) */
#define DGPMP2_FLAG_NONHOLONOMIC 1  /* planner_params['non_holonomic'];  needs dof == 3 */
#define DGPMP2_FLAG_VEL_LIMITS 2    /* planner_params['use_vel_limits']; needs dof == 2 */
#define DGPMP2_FLAG_Q_FULL 4        /* learn_params dynamics_mode == 'q_full': weights.qc_inv holds full d x d Q^-1 */
/* Fused learned-covariance head (replaces DiffGPMP2Planner.get_covariances, diff_gpmp2_planner.py:247-283):
 * the non-NULL pointers of dgpmp2_weights hold the RAW outputs of the learned module and the kernels form
 * the covariances themselves, with the reference's rounding (products taken in the I/O element type):
 *   w_obs = o*o (:275), eps = e*e (:279) and, per GP factor,
 *   qc_inv = q*q * I_dof from ONE value        ('diag_identity', :256-262; DGPMP2_FLAG_HEAD alone),
 *   qc_inv = v v^T from dof values             ('qc_full',  :269-273; with DGPMP2_FLAG_HEAD_QC_VEC),
 *   Q^-1   = v v^T from d = 2*dof values       ('q_full',   :274-278; with DGPMP2_FLAG_Q_FULL).
 * qc_stride_t is then the distance between the factors' raw values (1, dof or d for a packed vector).
 * The backward entry points still return the gradients w.r.t. the covariances (g_qc, g_w, g_eps). */
#define DGPMP2_FLAG_HEAD 8
#define DGPMP2_FLAG_HEAD_QC_VEC 16

/*
 * Constructor-time constants of the planner (reference plan_layer.py:14-85,
 * diff_gpmp2_planner.py:16-48).  All derived scalars are computed by the host in
 * double precision the way the reference computes them, so the kernels see the
 * same numbers as the reference's tensors.
 */
typedef struct dgpmp2_params {
  int32_t B;            /* problems in this call                                            */
  int32_t T;            /* num_traj_states = total_time_step + 1         (plan_layer.py:30) */
  int32_t dof;          /* 2 (PointRobot2D) or 3 (PointRobotXYH); state_dim d = 2*dof       */
  int32_t H, W;         /* SDF rows / cols; row 0 is y_hi                (sdf_utils.py:62)  */
  int32_t flags;        /* DGPMP2_FLAG_*                                                    */
  int64_t sdf_stride_b; /* elements between consecutive problems' SDFs (H*W; 0 = shared)    */
  double x_lo, y_lo;    /* env_params['x_lims'][0], ['y_lims'][0]                           */
  double res;           /* (x_hi - x_lo) / W  -- width only              (obstacle_cost.py:34) */
  double dt;            /* total_time_sec / total_time_step              (plan_layer.py:31) */
  double r_sphere;      /* robot sphere radius                           (obstacle_factor.py:37) */
  double ks_inv2;       /* 1 / K_s^2                                     (plan_layer.py:64) */
  double kg_inv2;       /* 1 / K_g^2                                     (plan_layer.py:65) */
  double reg;           /* optim_params['reg'] (delta)                   (plan_layer.py:96) */
  double kd_inv2;       /* 1 / K_d^2   (nonholonomic_factor.py:14)                          */
  double kv_inv2;       /* 1 / K_v^2   (velocity_limit_factor.py:15)                        */
  double vx_lim, vy_lim;/* gp_params['v_x'], ['v_y']                     (plan_layer.py:59-63) */
  double qc_inv[9];     /* static Qc^-1 (dof x dof, row-major) used when weights->qc_inv == NULL */
  double w_obs;         /* static 1/sigma_obs^2     used when weights->w_obs == NULL        */
  double eps;           /* static epsilon_dist      used when weights->eps   == NULL        */
  double qc_inv_fix[9]; /* constructor-time Qc^-1 for err_ext            (plan_layer.py:70-73) */
  double w_obs_fix;     /* constructor-time 1/sigma_obs^2 for err_ext    (plan_layer.py:74-76) */
} dgpmp2_params;

/*
 * Optional per-(problem, state) weights (the learned covariances of
 * diff_gpmp2_planner.py:183-205).  Element type follows the entry point's
 * suffix.  Any pointer may be NULL, in which case the static constant in
 * dgpmp2_params is used.  Strides are in ELEMENTS; a stride of 0 broadcasts.
 *   qc_inv : blocks of dof*dof (or d*d with DGPMP2_FLAG_Q_FULL), one per GP factor i in [0,T-1)
 *   w_obs  : one scalar per state (obscov_inv, nlinks == 1)
 *   eps    : one scalar per state
 * With DGPMP2_FLAG_HEAD the same pointers hold the raw head outputs instead (see the flag).
 */
typedef struct dgpmp2_weights {
  const void* qc_inv; int64_t qc_stride_b, qc_stride_t;
  const void* w_obs;  int64_t w_stride_b,  w_stride_t;
  const void* eps;    int64_t eps_stride_b, eps_stride_t;
} dgpmp2_weights;

int dgpmp2_abi_version(void);
const char* dgpmp2_status_string(int code);
/* Last CUDA error string seen by this thread (empty if none). */
const char* dgpmp2_last_cuda_error(void);

/*
 * One batched Gauss-Newton iteration -- replaces PlanLayer.forward
 * (plan_layer.py:87-99) = construct_linear_system_batch (:152-200) +
 * solve_linear_system_batch (:214-234) + error_batch (:273-308) +
 * error_ext_batch (:310-345), called from DiffGPMP2Planner.step
 * (diff_gpmp2_planner.py:176-211).
 *   th (B,T,d)  start, goal (B,d)  sdf (B,H,W)
 *   -> dth (B,T,d), err (B), err_ext (B)  [errors at the INPUT th, normalised by M]
 *   status (B) int32: 0 ok, t+1 if the factorisation met a non-positive pivot at state t
 *          (the reference's torch.cholesky would raise).  May be NULL.
 */
int dgpmp2_gn_step_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                       const float* sdf, const dgpmp2_weights* w,
                       float* dth, float* err, float* err_ext, int32_t* status, void* stream);
int dgpmp2_gn_step_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                       const double* sdf, const dgpmp2_weights* w,
                       double* dth, double* err, double* err_ext, int32_t* status, void* stream);

/*
 * dgpmp2_gn_step_f32 with one more output for tests / benchmarks: refine (B) int32 tells how each problem was
 * solved by the mixed-precision kernel (fp32 block cyclic reduction + fp64 residual refinement):
 *   k > 0  accepted after k refinement iterations,  k < 0  handed to the all-double path after |k| iterations,
 *   0      the launch used the all-double kernel for every problem (DGPMP2_PRECISION=64, or a shape the
 *          mixed-precision kernel does not take).
 */
int dgpmp2_gn_step_diag_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                            const float* sdf, const dgpmp2_weights* w,
                            float* dth, float* err, float* err_ext, int32_t* status, int32_t* refine, void* stream);

/*
 * Backward (reverse-mode derivative) of dgpmp2_gn_step_* -- replaces autograd through the
 * reference's dense A/b/K scatter, normal equations, Cholesky and inverses (plan_layer.py:152-234),
 * which is what makes the planner "differentiable" (learning/train_planner.py:366-374,
 * examples/diff_gpmp2_2d_example.py:75-78).  Recomputes the band from the inputs, solves
 * lambda = Lambda^-1 g_dth with the same block cyclic reduction and evaluates the factor VJPs.
 *   inputs : the forward inputs, dth = forward output (B,T,d), g_dth = dL/d dth (B,T,d),
 *            g_err_ext = dL/d err_ext (B) or NULL (err itself is not differentiable in the
 *            reference: it is computed under torch.no_grad, plan_layer.py:275)
 *   outputs (any may be NULL): g_th (B,T,d), g_start (B,d), g_goal (B,d),
 *            g_qc (B,T-1,dof,dof) -- or (B,T-1,d,d) with DGPMP2_FLAG_Q_FULL -- dense per-(b,t),
 *            g_w (B,T), g_eps (B,T),
 *            g_sdf (same layout as sdf; ACCUMULATED with atomics, the caller zero-fills it).
 */
int dgpmp2_gn_step_backward_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                                const float* sdf, const dgpmp2_weights* w, const float* dth, const float* g_dth,
                                const float* g_err_ext, float* g_th, float* g_start, float* g_goal, float* g_qc,
                                float* g_w, float* g_eps, float* g_sdf, void* stream);
int dgpmp2_gn_step_backward_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                                const double* sdf, const dgpmp2_weights* w, const double* dth, const double* g_dth,
                                const double* g_err_ext, double* g_th, double* g_start, double* g_goal, double* g_qc,
                                double* g_w, double* g_eps, double* g_sdf, void* stream);

/*
 * Optimise to convergence in ONE persistent launch -- replaces the per-sample
 * loop of DiffGPMP2Planner.forward (diff_gpmp2_planner.py:104-165) with static
 * or per-call-constant weights: per problem, th <- th + dth until
 * ||dth||_2 < tol_delta or iters >= max_iters (planner_utils.py:3-16).
 *   -> th_final (B,T,d), iters (B), err_per_iter / err_ext_per_iter (B,max_iters)
 *      [entry j = error at iterate j; entries >= iters[b] are left untouched],
 *      err_final (B) = error at th_final, status (B) as above.  Output pointers
 *      other than th_final and iters may be NULL.
 */
int dgpmp2_gn_solve_f32(const dgpmp2_params* p, const float* th_init, const float* start, const float* goal,
                        const float* sdf, const dgpmp2_weights* w, int32_t max_iters, double tol_delta,
                        float* th_final, int32_t* iters, float* err_per_iter, float* err_ext_per_iter,
                        float* err_final, float* err_ext_final, int32_t* status, void* stream);
int dgpmp2_gn_solve_f64(const dgpmp2_params* p, const double* th_init, const double* start, const double* goal,
                        const double* sdf, const dgpmp2_weights* w, int32_t max_iters, double tol_delta,
                        double* th_final, int32_t* iters, double* err_per_iter, double* err_ext_per_iter,
                        double* err_final, double* err_ext_final, int32_t* status, void* stream);

/*
 * Factor sweep without the solve -- replaces PlanLayer.error_batch (:273-308),
 * error_ext_batch (:310-345), start_goal_error / gp_error / obs_error (:374-388).
 * Outputs (each (B), any may be NULL): err, err_ext (weighted, / M);
 * err_sg = 0.5|e_s|^2 + 0.5|e_g|^2; err_gp = mean_i 0.5|g_i|^2; err_obs = mean_t 0.5 c_t^2.
 */
int dgpmp2_errors_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                      const float* sdf, const dgpmp2_weights* w,
                      float* err, float* err_ext, float* err_sg, float* err_gp, float* err_obs, void* stream);
int dgpmp2_errors_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                      const double* sdf, const dgpmp2_weights* w,
                      double* err, double* err_ext, double* err_sg, double* err_gp, double* err_obs, void* stream);

/*
 * Backward of dgpmp2_errors_* w.r.t. the trajectory -- replaces autograd through error_ext_batch (:310-345),
 * start_goal_error / gp_error / obs_error (:374-388), which the reference's training loss is built from
 * (learning/train_planner.py:327-346: unweighted_errors_batch(th_new) -> gp + sg + lambda * obs, error_ext).
 *   g_err_ext, g_err_sg, g_err_gp, g_err_obs: (B) upstream gradients, any may be NULL (= 0);
 *   err has no gradient (the reference computes it under torch.no_grad, :275).
 *   -> g_th (B,T,d) = sum_k g_err_k[b] * d err_k[b] / d th[b].  Only w->eps is read from the weights (the
 *   obstacle cost uses the installed eps; err_ext uses the constructor-time covariances).
 */
int dgpmp2_errors_backward_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                               const float* sdf, const dgpmp2_weights* w, const float* g_err_ext, const float* g_err_sg,
                               const float* g_err_gp, const float* g_err_obs, float* g_th, void* stream);
int dgpmp2_errors_backward_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                               const double* sdf, const dgpmp2_weights* w, const double* g_err_ext,
                               const double* g_err_sg, const double* g_err_gp, const double* g_err_obs, double* g_th,
                               void* stream);

/*
 * Stand-alone factor outputs -- replaces GPFactor.get_error (gp_factor.py:100-110),
 * ObstacleFactor.get_error (obstacle_factor.py:35-40) incl.
 * HingeLossObstacleCost.hinge_loss_signed_batch (obstacle_cost.py:29-38),
 * NonHolonomicFactor.get_error_full (nonholonomic_factor.py:16-30, applied per
 * problem) and VelocityLimitFactor.get_error_full (velocity_limit_factor.py:17-29).
 * Any output may be NULL.  Only w->eps is read from the weights.
 *   gp_err (B,T-1,d)   obs_cost (B,T)   obs_H (B,T,d)
 *   cust_err (B,T) [nonholonomic] or (B,T,2) [velocity limits]
 *   cust_H   (B,T,d)               or (B,T,2,d)
 */
int dgpmp2_factors_f32(const dgpmp2_params* p, const float* th, const float* sdf, const dgpmp2_weights* w,
                       float* gp_err, float* obs_cost, float* obs_H, float* cust_err, float* cust_H, void* stream);
int dgpmp2_factors_f64(const dgpmp2_params* p, const double* th, const double* sdf, const dgpmp2_weights* w,
                       double* gp_err, double* obs_cost, double* obs_H, double* cust_err, double* cust_H, void* stream);

/*
 * Batched bilinear SDF lookup with analytic gradient -- replaces
 * bilinear_interpolate (utils/sdf_utils.py:38-107).
 *   sdf (B,H,W) with problem stride sdf_stride_b, pts (B,N,2) -> dist (B,N), J (B,N,2)
 *   orig = (0 - lims[0]/res) per axis, computed by the caller as the reference does (:57-58).
 */
int dgpmp2_sdf_lookup_f32(const float* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_stride_b,
                          const float* pts, int32_t N, double res, double x_lo, double y_lo,
                          float* dist, float* J, void* stream);
int dgpmp2_sdf_lookup_f64(const double* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_stride_b,
                          const double* pts, int32_t N, double res, double x_lo, double y_lo,
                          double* dist, double* J, void* stream);

/*
 * Hinge obstacle cost of a batch of sphere centres -- replaces HingeLossObstacleCost.hinge_loss_signed_batch
 * (obstacle_cost.py:29-38) = bilinear_interpolate + the two torch.where selects, in one pass: per point 8 bytes of
 * position in, four SDF taps, cost (4 bytes) and the 2-wide gradient H_e = -J (8 bytes) out (fp32 I/O).
 *   pts (B,N,2) [2-element aligned], sdf (B,H,W) with problem stride sdf_stride_b (0 = shared)
 *   eps: per-point margin with element strides (eps_stride_b, eps_stride_n), or NULL -> eps_const;
 *   active when dist <= eps + r_sphere.   -> cost (B,N), He (B,N,2).
 */
int dgpmp2_hinge_batch_f32(const float* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_stride_b, const float* pts,
                           int32_t N, double res, double x_lo, double y_lo, const float* eps, int64_t eps_stride_b,
                           int64_t eps_stride_n, double eps_const, double r_sphere, float* cost, float* He, void* stream);
int dgpmp2_hinge_batch_f64(const double* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_stride_b, const double* pts,
                           int32_t N, double res, double x_lo, double y_lo, const double* eps, int64_t eps_stride_b,
                           int64_t eps_stride_n, double eps_const, double r_sphere, double* cost, double* He, void* stream);

/*
 * Signed distance field of a batch of occupancy images -- replaces sdf_2d (utils/sdf_utils.py:6-21,
 * datasets/utils.py:4-18: free = image > 0.75, optional padding with free cells, two
 * scipy.ndimage.distance_transform_edt calls, (edt(free) - edt(occupied)) * res).  Exact Euclidean
 * distance transform (integer squared distances, sqrt in double), including scipy's behaviour for
 * images without any background pixel.
 *   im (B,H,W) -> sdf_out (B, H+2*padlen, W+2*padlen); positive in free space.
 *   Limit: H+2p <= 254 and W+2p <= 254 (one-byte row distances), else DGPMP2_ERR_UNSUPPORTED.
 */
int dgpmp2_sdf_from_occupancy_f32(const float* im, int32_t B, int32_t H, int32_t W, int32_t padlen, double thresh,
                                  double res, float* sdf_out, void* stream);
int dgpmp2_sdf_from_occupancy_f64(const double* im, int32_t B, int32_t H, int32_t W, int32_t padlen, double thresh,
                                  double res, double* sdf_out, void* stream);
/* Same with a uint8 occupancy image (free = im > thresh, e.g. 0 / 255 images with thresh 191.25 = 0.75 * 255): a quarter of
 * the bytes of the float image when the maps come from the host. */
int dgpmp2_sdf_from_occupancy_u8_f32(const uint8_t* im, int32_t B, int32_t H, int32_t W, int32_t padlen, double thresh,
                                     double res, float* sdf_out, void* stream);

/* Same from BIT-PACKED occupancy maps (padlen 0): row-major, every row padded to 32-bit words, bit (x & 31) of word
 * x >> 5 set = pixel (y, x) is FREE.  1/32 of the bytes of the float SDF -- what makes shipping maps instead of SDFs
 * across PCIe pay (dgpmp2_gn_step_host_occ_f32 below). */
int dgpmp2_sdf_from_occupancy_bits_f32(const uint32_t* im_bits, int32_t B, int32_t H, int32_t W, double res, float* sdf_out,
                                       void* stream);

/*
 * The block-tridiagonal information system itself (what the reference holds as
 * dense A^T K A + reg I and A^T K b, plan_layer.py:217-220), always in double:
 *   D (B,T,d,d) diagonal blocks, U (B,T-1,d,d) blocks (t,t+1), r (B,T,d).
 * Inspection / parity entry point; the solve kernels never write the band to HBM.
 */
int dgpmp2_band_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                    const float* sdf, const dgpmp2_weights* w, double* D, double* U, double* r, void* stream);
int dgpmp2_band_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                    const double* sdf, const dgpmp2_weights* w, double* D, double* U, double* r, void* stream);

/*
 * End-to-end call with HOST buffers (th, start, goal, sdf and the outputs live in
 * host memory, pinned for full speed): copies the inputs to `dev_ws`, runs
 * dgpmp2_gn_step_*, copies dth / err / err_ext / status back, all on `stream`,
 * and synchronises the stream before returning.  `dev_ws` is a device buffer of
 * at least dgpmp2_host_step_workspace_bytes() bytes.  Weights must be static
 * (w == NULL).  `sdf_resident` says where this step's SDF is:
 *   DGPMP2_SDF_COPY (0)      `sdf` is copied to the workspace (and stays there for later calls);
 *   DGPMP2_SDF_RESIDENT (1)  the SDF already in the workspace (a previous DGPMP2_SDF_COPY call with the same shape) is
 *                            reused, `sdf` is ignored: Gauss-Newton iterations on fixed environments;
 *   DGPMP2_SDF_IN_PLACE (2)  `sdf` is read where it lies if it is pinned (mapped) host memory -- cudaHostAlloc /
 *                            cudaHostRegister / torch pin_memory(); dgpmp2_host_pointer_is_mapped() tells -- : the
 *                            kernel fetches only the 32-byte sectors its 4 taps per state touch over PCIe instead of the
 *                            whole field being copied first (B=1024, T=64, 128x128 fp32 maps on B200: 0.32 ms per step
 *                            instead of 1.42 ms, same bits).  Nothing is staged: the workspace SDF is left as it was.
 *                            Pageable `sdf`: falls back to DGPMP2_SDF_COPY.  The right mode for an SDF used once.
 * In every mode those of th / start / goal / dth / err / err_ext / status that are pinned (mapped) host memory are read
 * and written in place by the kernel instead of being copied (results are the same bits either way).
 * Opt-in (environment DGPMP2_HOST_CHUNKS=n > 1, per-problem SDFs, B >= 256): the batch is processed in n chunks that
 * alternate between two library-owned helper streams (created once per device, the only process-wide state of the
 * library; calls are serialised on them), so that one chunk's kernel and device-to-host copies overlap the next
 * chunk's host-to-device copy; the helper streams are ordered after `stream` on entry and `stream` after them on
 * exit.  Results are bit-identical to the one-launch path.  Off by default: the SDF copy saturates PCIe either way.
 */
#define DGPMP2_SDF_COPY 0
#define DGPMP2_SDF_RESIDENT 1
#define DGPMP2_SDF_IN_PLACE 2
int dgpmp2_host_step_workspace_bytes(const dgpmp2_params* p, int32_t elem_size, size_t* bytes);
/* 1 if `host_ptr` is pinned host memory addressable from the current device (what DGPMP2_SDF_IN_PLACE needs), else 0. */
int dgpmp2_host_pointer_is_mapped(const void* host_ptr);
int dgpmp2_gn_step_host_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                            const float* sdf, float* dth, float* err, float* err_ext, int32_t* status,
                            void* dev_ws, size_t dev_ws_bytes, int32_t sdf_resident, void* stream);
int dgpmp2_gn_step_host_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                            const double* sdf, double* dth, double* err, double* err_ext, int32_t* status,
                            void* dev_ws, size_t dev_ws_bytes, int32_t sdf_resident, void* stream);

/*
 * End-to-end call from HOST trajectories and HOST bit-packed occupancy maps (layout as
 * dgpmp2_sdf_from_occupancy_bits_f32; one H x W map per problem, p->sdf_stride_b == H*W): copies the inputs, builds the
 * signed distance fields on the device (exact EDT, = sdf_2d(map, padlen=0, res=p->res), generate_2d_dataset.py:211),
 * runs dgpmp2_gn_step_f32 and copies dth / err / err_ext / status back; synchronises the stream before returning.
 * Replaces the host-side sdf_2d + the B*H*W*4-byte SDF transfer of the reference's data path by B*H*ceil(W/32)*4 bytes.
 * `dev_ws`: at least dgpmp2_host_step_occ_workspace_bytes() bytes.
 */
int dgpmp2_host_step_occ_workspace_bytes(const dgpmp2_params* p, size_t* bytes);
int dgpmp2_gn_step_host_occ_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                                const uint32_t* occ_bits, float* dth, float* err, float* err_ext, int32_t* status,
                                void* dev_ws, size_t dev_ws_bytes, void* stream);

/*
 * Launch-shape query for benchmarking / tests: problems per CTA, threads per CTA
 * and dynamic shared-memory bytes the GN-step kernel will use for these params.
 * When an SM's share of the batch does not fit one CTA the grid holds CTAs of two sizes
 * (`problems_per_cta` in the first ones, one fewer in the later ones: balanced waves).
 */
int dgpmp2_gn_step_launch_shape(const dgpmp2_params* p, int32_t elem_size,
                                int32_t* problems_per_cta, int32_t* threads, int32_t* smem_bytes, int32_t* grid);

#ifdef __cplusplus
}
#endif
#endif /* DGPMP2_B200_H */
