/*
 * isv_capi.h -- C ABI of the B200-native IS-VINS marginalization + sparsification backend.
 *
 * This is the drop-in boundary.  The reference has no FFI: the path is
 * three argument-less C++ member functions that read/write `Estimator` members
 * (/root/reference/include/estimator.h:122-154).  Each entry point below cites the reference
 * code it replaces; the C++ shim in is_vins_b200/host/ (same class / method names as the
 * reference) packs the members into these PODs, so estimator.cpp keeps its call sites
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every matrix is COLUMN-MAJOR, i.e. exactly the memory of the Eigen::MatrixXd /
 *     Eigen::Matrix3d member it mirrors (`m.data()`), unless the field says row-major
 *     (the ceres `jacobians[i]` blocks are row-major, as ceres demands);
 *   - pose block  = [px,py,pz,qx,qy,qz,qw]          (src/estimator.cpp:474-487)
 *     speed-bias  = [v3, ba3, bg3]                   (src/estimator.cpp:489-499)
 *   - all arithmetic is IEEE FP64 on the GPU; there is NO CPU fallback: every compute entry
 *     point returns ISV_ERR_CUDA when no sm_100 device is usable;
 *   - no entry point allocates caller-visible memory, throws, or aborts; device scratch is owned
 *     by the handle; a handle is not thread-safe (one per host thread and GPU).
 */
#ifndef ISV_CAPI_H
#define ISV_CAPI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISV_ABI_VERSION 4   /* 2: isv_batch_in grew imu_raw / imu_init / imu_count / imu_k_max / flags (zero = v1 behaviour);
                               3: ... and lm_xy_f32 (NULL = v2 behaviour);
                               4: flags ISV_IN_TRI_RECORDS / ISV_OUT_TRI_RECORDS on the host-pointer entry point */

/* ---- status codes (SURVEY.md 8b "error conventions": the reference has none -- void + assert) */
typedef enum isv_status {
  ISV_OK = 0,
  ISV_ERR_BAD_ARG = 1,
  ISV_ERR_CUDA = 2,      /* no device / launch failure: loud, never a CPU fallback */
  ISV_ERR_ALLOC = 3
} isv_status;

/* per-window status bits written by the kernels (int32 bitmask, 0 == clean) */
#define ISV_W_NOT_SPD        0x01  /* an LLT met a non-positive pivot (reference: silent NaN)   */
#define ISV_W_RANK_DEFICIENT 0x02  /* fwd: FullPivHouseholderQR rank < 6 -> eigen path taken    */
#define ISV_W_NONFINITE      0x04  /* NaN/Inf in an output                                      */
#define ISV_W_NONUNIT_QUAT   0x08  /* | |q|^2 - 1 | > 1e-9 on an input pose (structured kernel
                                      assumes the unit quaternions ceres' Plus() guarantees)    */
#define ISV_W_EIG_NOCONV     0x10  /* Jacobi eigensolver hit its sweep cap                      */
#define ISV_W_SINGULAR       0x20  /* an LU inverse met a zero pivot                            */

/* ---- configuration: the globals the reference reads inside Marg*  ------ */
typedef struct isv_config {
  double alpha;              /* ALPHA, config/euroc_config.yaml:86 ; eigenvalue kept iff > alpha */
  double proj_sqrt_info[4];  /* ProjectionFactor::sqrt_info 2x2, src/estimator.cpp:35           */
  double g[3];               /* G, src/parameters.cpp:19,96                                      */
  double acc_n, gyr_n, acc_w, gyr_w; /* IntegrationBase noise, integration_base.h:21-27          */
  int vo_size;               /* Vo_SIZE, include/parameters.h:35                                 */
  int all_buf_size;          /* ALL_BUF_SIZE, include/parameters.h:40                            */
  int qr_rank_eps_log10;     /* -16: `eps` of src/estimator.cpp:8 used at :1305                  */
  int reserved;
} isv_config;

typedef struct isv_handle isv_handle;

/* fills *cfg with config/euroc_config.yaml's values */
void isv_default_config(isv_config* cfg);
int isv_abi_version(void);
/* human-readable name of a status code */
const char* isv_status_string(isv_status s);

isv_status isv_create(const isv_config* cfg, int device, isv_handle** out);
void isv_destroy(isv_handle* h);
/* the CUDA stream every call of this handle is ordered on (cudaStream_t as void*) */
void* isv_stream(isv_handle* h);
/* use an external stream (e.g. torch's current stream); 0 restores the handle's own stream */
isv_status isv_set_stream(isv_handle* h, void* cuda_stream);
isv_status isv_synchronize(isv_handle* h);
/* number of kernel launches issued by this handle so far (bench.py's gpu_launches) */
int64_t isv_launch_count(const isv_handle* h);
/* Tuning knobs (A/B measurements, tests).  Environment variables of the same meaning seed them at isv_create:
 *   ISV_TUNE_FUSED_MAX_WINDOWS (env ISV_FUSED_MAX, default 148): isv_marg_window_batch(ISV_RUN_BOTH) calls of at most this
 *       many windows take the one-launch fused kernel (one CTA per window); 0 = always the warp-per-window batch kernels.
 *   ISV_TUNE_EVENT_MODE (env ISV_EVENT_MODE, default 0): route of isv_marg_event -- 0 zero-copy fused kernel, 1 fused
 *       kernel on a device-side mirror (H2D, launch, D2H, stream synchronise), 2 batch kernels on the mirror.
 *   ISV_TUNE_ACC_PERSIST (env ISV_ACC_PERSIST, default 0 = one CTA per window): persistent warps per SM of the landmark
 *       kernel for batches of >= 2368 windows -- an occupancy experiment (leave registers to the backward kernel so that
 *       it runs beside the landmark phase), measured without gain and kept for re-measurement.                         */
#define ISV_TUNE_FUSED_MAX_WINDOWS 1
#define ISV_TUNE_EVENT_MODE 2
#define ISV_TUNE_ACC_PERSIST 3
isv_status isv_set_tuning(isv_handle* h, int knob, int value);

/* ---- index maps: the bit-exact contract -------------------------------
 * Each writes (offset, dim) per block into out[2*i], out[2*i+1] and returns the number of blocks.
 *   init  (src/estimator.cpp:747-758): T0..T_{V-1}, VB_{V-1}, VB_0..VB_{V-2}       -> 2V blocks
 *   fwd   (src/estimator.cpp:1153-1162): T1, T0, landmark 0..L-1                   -> 2+L blocks
 *   bwd   (src/estimator.cpp:1358-1366): T_V, VB_V, T_{V-1}, VB_{V-1}              -> 4 blocks  */
int isv_order_map_init(int vo_size, int32_t* out);
int isv_order_map_forward(int n_landmarks, int32_t* out);
int isv_order_map_backward(int vo_size, int32_t* out);

/* ---- record sizes (doubles) ------------------------------------------------------------------ */
#define ISV_POSE 7
#define ISV_SB 9
#define ISV_SE3_REC 48      /* t[3], R[9], sqrt_info[36]            (SE3PriorFactor members)      */
#define ISV_REL_REC 48      /* delta_t[3], delta_R[9], sqrt_info[36] (RelativePoseFactor members)  */
#define ISV_VB_REC 90       /* VB[9], sqrt_info[81]                  (Linear9Factor members)       */
#define ISV_RP_IN_REC 5     /* valid (0/1), sqrt_info[4]  : vioRollPitchEdges[0] if index==0      */
#define ISV_RP_REC 13       /* R[9], sqrt_info[4]                    (RollPitchFactor members)     */
#define ISV_PG_REC 89       /* delta_t[3], delta_R[9], sqrt_info[36], covRel[36], distance, covAbs[4]
                               (CombinedFactors, include/factor/pose_graph_factors.h:6-17)        */
#define ISV_PREINT_REC 467  /* delta_p[3], delta_q[4] (x,y,z,w), delta_v[3], linearized_ba[3],
                               linearized_bg[3], sum_dt, jacobian[225], covariance[225]
                               (IntegrationBase members, integration_base.h:188-203)              */
#define ISV_LM_COMPONENTS 6 /* x_i, y_i, z_i, x_j, y_j, inv_dep                                    */

/* ---- batched MargForward + MargBackward over independent windows (device pointers) ----------
 * One "window" = one MARGIN_OLD event = Estimator::MargForward (src/estimator.cpp:1149-1352)
 * followed by Estimator::MargBackward (:1354-1539).  All pointers are DEVICE pointers, SoA over
 * windows; landmark observations are component-major and CSR-indexed by lm_offset so that a warp
 * reads 32 consecutive landmarks of one component in one coalesced 256-byte request.           */
typedef struct isv_batch_in {
  int32_t n_windows;
  int32_t ex_pose_shared;        /* 1: ex_pose is one [7] record for all windows               */
  const int64_t* lm_offset;      /* [n+1] first landmark of each window in lm_obs               */
  const double* lm_obs;          /* [6][lm_stride]: pts_i.x, pts_i.y, pts_i.z, pts_j.x, pts_j.y,
                                    inv_dep  (forwardProjectiontoSparsify[k]->pts_i/pts_j,
                                    para_Feature[MargPointIdx[k]])                              */
  int64_t lm_stride;             /* doubles between components (>= lm_offset[n])                */
  const double* pose_fwd;        /* [n][2][7] para_Pose[0], para_Pose[1]                        */
  const double* ex_pose;         /* [n][7] or [7]  para_Ex_Pose[0]                              */
  const double* prior_se3;       /* [n][48]  vioPosePriorEdge            (t, R, sqrt_info)      */
  const double* prior_rel;       /* [n][48]  vioRelativePoseEdges[1]     (delta_t, delta_R, s)  */
  const double* prior_rp;        /* [n][5]   vioRollPitchEdges[0] (may be NULL: none valid)     */
  const double* pose_bwd;        /* [n][2][7] para_Pose[V-1], para_Pose[V]                      */
  const double* sb_bwd;          /* [n][2][9] para_SpeedBias[V-1], para_SpeedBias[V]            */
  const double* prior_vb;        /* [n][90]  vioVBPrior                                         */
  const double* preint;          /* [n][467] backwardIMUtoSparsify->pre_integration; NULL: rebuilt
                                    on the GPU from imu_raw / imu_init below                    */
  /* ---- ABI 2 (all zero = ABI 1 behaviour) --------------------------------------------------
   * The pre-integration record is a pure function of the interval's raw IMU samples
   * (IntegrationBase::push_back, include/factor/integration_base.h:30-158): a caller on the far
   * side of PCIe ships 12 + 7 K doubles instead of 467 and the batch call runs
   * preintegrate_kernel first.                                                                 */
  const double* imu_raw;         /* [n][imu_k_max][7] dt, acc[3], gyr[3] (dt_buf/acc_buf/gyr_buf) */
  const double* imu_init;        /* [n][12] acc_0, gyr_0, linearized_ba, linearized_bg          */
  const int32_t* imu_count;      /* [n] samples used per window, or NULL = imu_k_max everywhere */
  int32_t imu_k_max;             /* samples stored per window                                   */
  int32_t flags;                 /* ISV_IN_* bits                                               */
  /* ---- ABI 3 (NULL = ABI 2 behaviour) ---------------------------------------------------------
   * pts_i.x / pts_i.y are FP32 values in the reference: the feature tracker stores the undistorted
   * points as cv::Point2f (include/feature_tracker/feature_tracker_simple.h:55,
   * src/feature_tracker/feature_tracker_simple.cpp:207) and System widens them to double
   * (src/System.cpp:119-122).  A caller that still has the floats hands them over as they are:
   * components 0 and 1 of lm_obs are then neither read nor copied, the kernels widen on load
   * (exact), and a landmark costs 4 + 4 + 8 bytes of PCIe / HBM traffic instead of 24.           */
  const float* lm_xy_f32;        /* [2][lm_stride] pts_i.x, pts_i.y as float, or NULL            */
} isv_batch_in;

/* pts_i.z == 1 for every landmark (the feature tracker normalises: src/System.cpp:346): component 2 of
 * lm_obs is neither read by the kernels nor copied by the host entry point (which spot-checks it). */
#define ISV_IN_PTS_I_Z_ONE 1
/* ABI 4, host-pointer entry point only (isv_marg_window_batch_host): the records cross PCIe WITHOUT their structural zeros.
 * Every sqrt_info in the path is upper triangular (it is an `LLT(...).matrixL().transpose()`, src/estimator.cpp:1349,
 * :1500-1516, include/factor/*.h) and covRel is symmetric, so a 6 x 6 block is 21 numbers, a 9 x 9 block 45, a 2 x 2 block 3.
 *   ISV_IN_TRI_RECORDS   the prior records arrive packed: prior_se3 / prior_rel [n][33] = t | delta_t 3, R 9 (column-major),
 *                        upper triangle of sqrt_info column by column 21; prior_vb [n][54] = VB 9 + 45; prior_rp [n][4] =
 *                        valid + 3.  (319 -> 252 doubles per window together with the other records.)
 *   ISV_OUT_TRI_RECORDS  the results are delivered packed: se3_out / rel_out [n][33]; pg_out [n][59] = delta_t 3, delta_R 9,
 *                        sqrt_info 21, covRel (upper triangle) 21, distance, covAbs 4; vb_out [n][54]; rp_out [n][12] =
 *                        R 9 + 3.  (289 -> 191 doubles per window.)
 * The kernels keep working on the full records in HBM; two small kernels expand / compact them on the device.        */
#define ISV_IN_TRI_RECORDS 2
#define ISV_OUT_TRI_RECORDS 4
#define ISV_SE3_TRI_REC 33
#define ISV_REL_TRI_REC 33
#define ISV_VB_TRI_REC 54
#define ISV_RP_IN_TRI_REC 4
#define ISV_PG_TRI_REC 59
#define ISV_RP_TRI_REC 12

typedef struct isv_batch_out {
  double* se3_out;               /* [n][48]  forwardPosePriorEdgeToAdd                          */
  double* pg_out;                /* [n][89]  CombinedFactors pushed to pose_graph_factors_buf   */
  double* rel_out;               /* [n][48]  backwardRelativePoseEdgeToAdd                      */
  double* vb_out;                /* [n][90]  backwardVBEdgeToAdd                                */
  double* rp_out;                /* [n][13]  rollPitchFactor pushed to vioRollPitchEdges        */
  int32_t* rank;                 /* [n][2]   fwd: QR rank of Lamda_prior (eigen rank if <6);
                                             bwd: #eigenvalues > alpha                          */
  int32_t* status;               /* [n]      ISV_W_* bitmask                                    */
} isv_batch_out;

#define ISV_RUN_FORWARD  1
#define ISV_RUN_BACKWARD 2
#define ISV_RUN_BOTH     3
/* profiling aids: the two stages of MargForward on their own.  STAGE1 = landmark phase (Jacobians +
 * Gram accumulation, leaves its result in the handle's scratch), STAGE2 = the 12x12 / 6x6 tail that
 * consumes the scratch of the preceding STAGE1 call on the same handle and batch.                */
#define ISV_RUN_FORWARD_STAGE1 4
#define ISV_RUN_FORWARD_STAGE2 8
/* likewise: the factor-Jacobian launches alone (thread per window and factor; they feed STAGE2 and the
 * backward kernel through the handle's scratch), and the backward kernel alone                    */
#define ISV_RUN_FACTOR_JAC 16
#define ISV_RUN_BACKWARD_STAGE2 32

/* stream-ordered, asynchronous; `which` selects MargForward / MargBackward / both.
 * Unused inputs/outputs of a skipped half may be NULL.
 * Routes: ISV_RUN_BOTH on at most ISV_TUNE_FUSED_MAX_WINDOWS (148) windows with `preint` given takes ONE launch of the fused
 * kernel (one CTA per window, is_vins_b200/csrc/isv_event_kernel.cuh); everything else the warp-per-window batch kernels
 * (7 launches forked over three streams; replayed as one cached CUDA graph for repeated calls on up to 4736 windows).
 * Both routes run the same device functions; results agree to rounding (the fused kernel sums the landmark Gram as seven
 * partial sums).                                                                                  */
isv_status isv_marg_window_batch(isv_handle* h, const isv_batch_in* in, const isv_batch_out* out,
                                 int which);

/* same call with HOST pointers in both structs: copies the per-window records once, then pipelines the landmark
 * components (H2D, one ordered stream), the kernels and the results (D2H) in four chunks on four streams, and
 * synchronises.  Pinned host buffers make the copies asynchronous.  This is what the e2e number in bench.py times.   */
isv_status isv_marg_window_batch_host(isv_handle* h, const isv_batch_in* in,
                                      const isv_batch_out* out, int which);

/* ---- forensic mode (NOT the hot path; SURVEY.md 7.3 item 5: the dense route "must be available") --------------
 * isv_marg_forensic_batch = isv_marg_window_batch(ISV_RUN_BOTH) plus:
 *   - the structured route's intermediates: Lamda_prior of MargForward (src/estimator.cpp:1288) and the factor G with
 *     G^T G = Lamda_prior of MargBackward (:1419), so that a parity break can be localised on the GPU;
 *   - the reference's KLD diagnostics (computed and discarded there): forward :1333-1345, backward :1522-1534 together
 *     with the absolute-position / yaw informations of :1518-1519 that only feed it.
 * All pointers are DEVICE pointers; stream-ordered.                                                              */
typedef struct isv_forensic_out {
  double* lamda_prior_fwd;       /* [n][36]  6x6 column-major (required)                                */
  double* g_bwd;                 /* [n][315] 15x21, row k at [21 k + c] (required)                       */
  double* kld_fwd;               /* [n] NaN when FullPivHouseholderQR's rank < 6 (required)              */
  double* kld_bwd;               /* [n] (required)                                                       */
  double* lamda_prior_bwd;       /* [n][441] 21x21 = G^T G, column-major (may be NULL)                   */
  double* eig_bwd;               /* [n][21] eigenvalues of it, ascending (may be NULL)                   */
  double* info_abs;              /* [n][9]  (may be NULL)                                                */
  double* info_yaw;              /* [n]     (may be NULL)                                                */
} isv_forensic_out;
isv_status isv_marg_forensic_batch(isv_handle* h, const isv_batch_in* in, const isv_batch_out* out,
                                   const isv_forensic_out* dbg);
/* The reference's dense route on normal equations that are already assembled (isv_build_normal_equations):
 * A [n_problems][n*n] column-major symmetric, kept block [0, m0), marginalized block [m0, n).  Inverts the WHOLE
 * marginalized block with a full-pivot elimination (the algorithm class of `fullPivLu().solve(Identity)`,
 * src/estimator.cpp:1286, :1417; dependent unknowns zero-filled) and returns A_rr - A_rm A_mm^-1 A_rm^T
 * (A_prior [n_problems][m0*m0]), optionally the inverse (Amm_inv [n_problems][m*m], may be NULL) and the number of
 * accepted pivots (rank, may be NULL).  n - m0 <= 1024.  Device pointers, stream-ordered, O(m^3) per problem.  */
isv_status isv_literal_schur(isv_handle* h, int n_problems, int n, int m0, const double* A, double* A_prior,
                             double* Amm_inv, int32_t* rank);

/* ---- single-window convenience wrappers (host pointers, blocking) ----------------------------
 * These are what the shim's Estimator::MargForward()/MargBackward() call.                       */
typedef struct isv_fwd_in {
  int32_t n_landmarks;           /* MargPointIdx.size()                                          */
  const double* pose0;           /* para_Pose[0]                                                 */
  const double* pose1;           /* para_Pose[1]                                                 */
  const double* ex_pose;         /* para_Ex_Pose[0]                                              */
  const double* inv_dep;         /* [L]    para_Feature[MargPointIdx[k]][0]                      */
  const double* pts_i;           /* [L][3] forwardProjectiontoSparsify[k]->pts_i                 */
  const double* pts_j;           /* [L][3] forwardProjectiontoSparsify[k]->pts_j                 */
  const double* prior_se3;       /* [48]   vioPosePriorEdge                                      */
  const double* prior_rel;       /* [48]   vioRelativePoseEdges[1]                               */
  const double* prior_rp;        /* [5] or NULL                                                  */
} isv_fwd_in;

typedef struct isv_fwd_out {
  double se3[ISV_SE3_REC];
  double pg[ISV_PG_REC];
  int32_t rank;
  int32_t status;
} isv_fwd_out;

typedef struct isv_bwd_in {
  const double* pose_i;          /* para_Pose[V-1]                                               */
  const double* sb_i;            /* para_SpeedBias[V-1]                                          */
  const double* pose_j;          /* para_Pose[V]                                                 */
  const double* sb_j;            /* para_SpeedBias[V]                                            */
  const double* prior_vb;        /* [90]                                                         */
  const double* preint;          /* [467]                                                        */
} isv_bwd_in;

typedef struct isv_bwd_out {
  double rel[ISV_REL_REC];
  double vb[ISV_VB_REC];
  double rp[ISV_RP_REC];
  int32_t rank;
  int32_t status;
} isv_bwd_out;

isv_status isv_marg_forward(isv_handle* h, const isv_fwd_in* in, isv_fwd_out* out);
isv_status isv_marg_backward(isv_handle* h, const isv_bwd_in* in, isv_bwd_out* out);
/* One whole MARGIN_OLD event -- `MargForward(); MargBackward();` as Estimator::backendOptimization() calls them back to
 * back (src/estimator.cpp:1555-1558; the two read disjoint members and neither reads the other's outputs) -- in ONE
 * blocking call and ONE kernel launch (marg_event_fused_kernel: a CTA of eight warps owns the event; the factor-Jacobian
 * chains, the landmark phase split over seven warps, the forward tail and the backward chain run side by side).  By default
 * the event is packed into a mapped pinned block that the kernel reads and writes over PCIe (no copy engine, no stream
 * synchronisation: the host spins on a completion word) -- ISV_TUNE_EVENT_MODE selects the staged routes instead.
 * Results equal the two separate calls to rounding (<= 1e-12 relative: the landmark Gram is summed as seven partial sums);
 * with ISV_TUNE_EVENT_MODE = 2 they are bit-identical.  Both `status` fields carry the OR of the event's warnings.  */
isv_status isv_marg_event(isv_handle* h, const isv_fwd_in* fwd_in, const isv_bwd_in* bwd_in, isv_fwd_out* fwd_out,
                          isv_bwd_out* bwd_out);

/* ---- initFactorGraph sparsification tail (src/estimator.cpp:745-1001) ---------------------------
 * One-time: the V-1 IMU factors of the initial window -> V-1 RelativePoseFactors
 * (vioRelativePoseEdges[1..V-1]), the SE3PriorFactor on T_0 (vioPosePriorEdge) and the Linear9Factor
 * on VB_{V-1} (vioVBPrior).  V = isv_config.vo_size (2..10).  SoA over independent windows.        */
typedef struct isv_init_in {
  int32_t n_windows;
  const double* poses;           /* [n][V][7]     para_Pose[0..V-1]                              */
  const double* sbs;             /* [n][V][9]     para_SpeedBias[0..V-1]                         */
  const double* preint;          /* [n][V-1][467] pre_integrations[i+1] (links frame i -> i+1)   */
} isv_init_in;

typedef struct isv_init_out {
  double* rel_out;               /* [n][V-1][48]  vioRelativePoseEdges[i+1]                      */
  double* se3_out;               /* [n][48]       vioPosePriorEdge                               */
  double* vb_out;                /* [n][90]       vioVBPrior                                     */
  int32_t* rank;                 /* [n]           #eigenvalues of the 6V+9 marginal > alpha      */
  int32_t* status;               /* [n]           ISV_W_* bitmask (may be NULL)                  */
} isv_init_out;

/* device pointers, stream-ordered */
isv_status isv_init_sparsify_batch(isv_handle* h, const isv_init_in* in, const isv_init_out* out);
/* host pointers, blocking */
isv_status isv_init_sparsify_host(isv_handle* h, const isv_init_in* in, const isv_init_out* out);

/* ---- IMU pre-integration (include/factor/integration_base.h:30-36, :54-158) -----------------------
 * IntegrationBase(acc_0, gyr_0, linearized_ba, linearized_bg) followed by push_back(dt, acc, gyr) for
 * every sample: midpoint rule + jacobian / covariance propagation with the handle's noise densities.
 * preint_out is the 467-double record MargBackward reads (ISV_PREINT_REC).                        */
typedef struct isv_preint_in {
  int32_t n;                     /* intervals                                                    */
  int32_t k_max;                 /* samples stored per interval                                  */
  const int32_t* k_count;        /* [n] samples used per interval, or NULL = k_max everywhere    */
  const double* imu_raw;         /* [n][k_max][7] dt, acc[3], gyr[3]   (dt_buf/acc_buf/gyr_buf)  */
  const double* imu_init;        /* [n][12] acc_0, gyr_0, linearized_ba, linearized_bg           */
} isv_preint_in;

/* device pointers, stream-ordered */
isv_status isv_preintegrate_batch(isv_handle* h, const isv_preint_in* in, double* preint_out);
/* host pointers, blocking */
isv_status isv_preintegrate_host(isv_handle* h, const isv_preint_in* in, double* preint_out);

/* ---- ceres CostFunction::Evaluate contract, batched (SURVEY.md 8a rows 2,3,5,7-10; 8f rank 1) -----
 * The factors problemSolve() hands to ceres (src/estimator.cpp:1004-1146).  For every factor the
 * kernels write what `Evaluate(parameters, residuals, jacobians)` writes: residuals = sqrt_info * r
 * and, per parameter block, jacobians[i] = sqrt_info * d r / d block_i as a ROW-MAJOR
 * num_residuals x global_size matrix whose 7th pose column is zero (the tangent convention of
 * PoseLocalParameterization, src/factor/pose_local_parameterization.cpp:3-27).  A NULL output
 * pointer is ceres' `jacobians == nullptr` / `jacobians[i] == nullptr` for that block.
 * Parameter blocks are gathered by index; the blocks of many windows may be concatenated.       */
typedef struct isv_param_blocks {
  int32_t n_pose, n_speed_bias, n_ex_pose, n_feature;
  const double* pose;            /* [n_pose][7]        para_Pose        include/estimator.h:122    */
  const double* speed_bias;      /* [n_speed_bias][9]  para_SpeedBias   :123                       */
  const double* ex_pose;         /* [n_ex_pose][7]     para_Ex_Pose     :125                       */
  const double* feature;         /* [n_feature]        para_Feature     :124 (inverse depth)       */
} isv_param_blocks;

#define ISV_W_BAD_INDEX      0x40  /* a factor referenced a parameter block out of range: skipped  */
#define ISV_W_DIAG_COUPLED   0x80  /* two DIFFERENT scalars of the diagonal marginalized block [m_dense, m_dense + m_diag)
                                      are coupled by a factor (or by the previous prior): the diagonal elimination would
                                      silently drop that coupling -- put such blocks into the dense block instead      */

/* ProjectionFactor::Evaluate  src/factor/projection_factor.cpp:24-122 ; sqrt_info = isv_config    */
typedef struct isv_proj_factors {
  int64_t n;                     /* factors                                                      */
  int64_t stride;                /* elements between components of idx / obs (>= n)              */
  const int32_t* idx;            /* [4][stride] imu_i, imu_j, ex pose index, feature_index
                                    (ProjectionFactor::setIndex, src/estimator.cpp:1081 ; the
                                    AddResidualBlock argument order :1090)                       */
  const double* obs;             /* [5][stride] pts_i.x, pts_i.y, pts_i.z, pts_j.x, pts_j.y      */
  double cauchy_a;               /* 0: no loss.  > 0: ceres::CauchyLoss(a) applied the way ceres'
                                    Corrector does for rho'' <= 0: r, J scaled by sqrt(rho'(|r|^2))
                                    (loss_function of :1018 has a = 1)                           */
  /* ProjectionTdFactor (online time-offset estimation; north_star / BASELINE configs[3]).  ABSENT from
   * the reference (only `para_Td` and the per-observation velocity / td fields
   * exist, include/estimator.h:126, include/feature_tracker/feature_manager.h:24-39): follows VINS-Mono's
   * published projection_td_factor.cpp, parity unpinned.  td_obs == NULL -> plain ProjectionFactor.  */
  const double* td_obs;          /* [8][stride] velocity_i.x, velocity_i.y, velocity_j.x, velocity_j.y,
                                    td_i, td_j, row_i - ROW/2, row_j - ROW/2                      */
  const double* td;              /* [n_td] para_Td[.][0]                                          */
  const int32_t* td_idx;         /* [n] index into td, or NULL = 0                                */
  int32_t n_td;
  int32_t reserved;
  double tr_over_row;            /* TR / ROW (rolling-shutter read-out per image row), 0 = global shutter */
} isv_proj_factors;
typedef struct isv_proj_eval {
  double* residuals;             /* [n][2]                                                       */
  double* jac_pose_i;            /* [n][2][7] row-major, or NULL                                 */
  double* jac_pose_j;            /* [n][2][7] or NULL                                            */
  double* jac_ex_pose;           /* [n][2][7] or NULL (SetParameterBlockConstant, :1037)         */
  double* jac_feature;           /* [n][2]    or NULL                                            */
  double* jac_td;                /* [n][2]    or NULL (ProjectionTdFactor's 5th block)            */
} isv_proj_eval;

/* IMUFactor::Evaluate  include/factor/imu_factor.h:23-159 (+ integration_base.h:160-186);
 * sqrt_info = LLT(covariance^-1).matrixL().transpose() is recomputed per call as the reference does */
#define ISV_IMU_JAC_REC 480      /* 15x7 | 15x9 | 15x7 | 15x9 row-major at offsets 0,105,240,345 */
typedef struct isv_imu_factors {
  int32_t n;
  const int32_t* idx;            /* [n][2] imu_i, imu_j (pose and speed-bias share the frame index) */
  const double* preint;          /* [n][467] pre_integrations[j] (ISV_PREINT_REC)                */
} isv_imu_factors;
typedef struct isv_imu_eval {
  double* residuals;             /* [n][15]                                                      */
  double* jacobians;             /* [n][480] or NULL                                             */
} isv_imu_eval;

/* the recovered / prior factors: RelativePoseFactor (relative_pose_factor.h:27-70), SE3PriorFactor
 * (se3_prior_factor.h:21-51), Linear9Factor (linear9_factor.h:20-44), RollPitchFactor
 * (rollpitch_factor.h:26-57), YawFactor (yaw_factor.h:23-49).  Records as in the Marg* outputs.   */
#define ISV_YAW_REC 4            /* yaw_meas[3], sqrt_info[1]                                    */
typedef struct isv_small_factors {
  int32_t n_rel, n_se3, n_vb, n_rp, n_yaw;
  const int32_t* rel_idx;        /* [n_rel][2] imu_i, imu_j                                      */
  const double* rel_rec;         /* [n_rel][48]                                                  */
  const int32_t* se3_idx;        /* [n_se3]                                                      */
  const double* se3_rec;         /* [n_se3][48]                                                  */
  const int32_t* vb_idx;         /* [n_vb]  speed-bias block index                               */
  const double* vb_rec;          /* [n_vb][90]                                                   */
  const int32_t* rp_idx;         /* [n_rp]                                                       */
  const double* rp_rec;          /* [n_rp][13]                                                   */
  const int32_t* yaw_idx;        /* [n_yaw]                                                      */
  const double* yaw_rec;         /* [n_yaw][4]                                                   */
  double cauchy_a;               /* as isv_proj_factors.cauchy_a (problemSolve uses a = 1, :1106-1121) */
} isv_small_factors;
typedef struct isv_small_eval {
  double* rel_res;  double* rel_jac;   /* [n_rel][6],  [n_rel][84]  (6x7 | 6x7) or NULL           */
  double* se3_res;  double* se3_jac;   /* [n_se3][6],  [n_se3][42]                                */
  double* vb_res;   double* vb_jac;    /* [n_vb][9],   [n_vb][81]                                 */
  double* rp_res;   double* rp_jac;    /* [n_rp][2],   [n_rp][14]                                 */
  double* yaw_res;  double* yaw_jac;   /* [n_yaw][1],  [n_yaw][7]                                 */
} isv_small_eval;

/* device pointers, stream-ordered.  `status` (device int32, may be NULL) receives ISV_W_* bits.   */
isv_status isv_eval_projection_batch(isv_handle* h, const isv_param_blocks* pb, const isv_proj_factors* f,
                                     const isv_proj_eval* out, int32_t* status);
isv_status isv_eval_imu_batch(isv_handle* h, const isv_param_blocks* pb, const isv_imu_factors* f,
                              const isv_imu_eval* out, int32_t* status);
isv_status isv_eval_small_batch(isv_handle* h, const isv_param_blocks* pb, const isv_small_factors* f,
                                const isv_small_eval* out, int32_t* status);

/* One problemSolve() iteration: every factor class in one call (any of the three factor sets, with
 * its output struct, may be NULL).  The projection kernel runs on the handle's stream; the IMU and
 * prior-factor kernels are forked onto internal streams and joined back, so the call is still
 * stream-ordered as a whole.  This is what a ceres::EvaluationCallback binds.    */
isv_status isv_eval_problem(isv_handle* h, const isv_param_blocks* pb, const isv_proj_factors* pf,
                            const isv_proj_eval* po, const isv_imu_factors* mf, const isv_imu_eval* mo,
                            const isv_small_factors* sf, const isv_small_eval* so, int32_t* status);

/* ---- device-resident factor state of n independent sequences ------------
 * The members Estimator carries from frame to frame -- vioRelativePoseEdges[1..V-1], vioPosePriorEdge,
 * vioVBPrior, vioRollPitchEdges (include/estimator.h:137-154) and the pose-graph accumulator
 * `accumFactor` (src/pose_graph/pose_graph_builder.cpp:157) -- stay in HBM; per frame the host sends
 * only states and observations.  One call per reference step:
 *   isv_seq_init         initFactorGraph's sparsification tail + factor installation  :745-1001
 *   isv_seq_update       factor->update(...) after problemSolve                       :1133-1144
 *   isv_seq_yaw          double2vector()'s rotation of the two priors                 :520-550
 *   isv_seq_marginalize  MargForward + MargBackward reading the priors from the state (:1149-1539),
 *                        CombinedFactors::operator+ / keyframe cut (pose_graph_factors.h:27-51,
 *                        pose_graph_builder.cpp:157-158,214), slideWindow()'s factor rotation (:1605-1638)
 * All pointers are DEVICE pointers unless a name says host; calls are stream-ordered.              */
#define ISV_ACC_REC 119        /* CombinedFactors as a POD record:                                  */
#define ISV_ACC_COVREL 48      /*   delta_t[3] delta_R[9] sqrt_info[36] | covRel[36]                */
#define ISV_ACC_DISTANCE 84    /*   distance, length, vio_index, pg_index, ts                       */
#define ISV_ACC_LENGTH 85
#define ISV_ACC_VIO_INDEX 86
#define ISV_ACC_PG_INDEX 87
#define ISV_ACC_TS 88
#define ISV_ACC_RI 89          /*   Ri[9] (column-major), ti[3]                                     */
#define ISV_ACC_TI 98
#define ISV_ACC_RP_VALID 101   /*   rollPitchFactor != nullptr, its record R[9] sqrt_info[4]        */
#define ISV_ACC_RP 102
#define ISV_ACC_COVABS 115     /*   covAbs[4] of the last factor added (operator+ itself drops it)  */

typedef struct isv_seq isv_seq;
isv_status isv_seq_create(isv_handle* h, int n_sequences, isv_seq** out);
void isv_seq_destroy(isv_handle* h, isv_seq* s);
/* rank: device int32 [n] (#eigenvalues > alpha of the init marginal), may be NULL */
isv_status isv_seq_init(isv_handle* h, isv_seq* s, const isv_init_in* in, int32_t* rank);

typedef struct isv_seq_update_in {
  const double* old_P;           /* [n][V][3]  Ps[0..V-1] before the solve                          */
  const double* old_R;           /* [n][V][9]  Rs[0..V-1] before the solve (column-major)           */
  const double* old_vb;          /* [n][9]     Vs, Bas, Bgs of frame V-1 before the solve           */
  const double* pose;            /* [n][V][7]  para_Pose[0..V-1] after the solve                    */
  const double* speed_bias;      /* [n][9]     para_SpeedBias[V-1] after the solve                  */
} isv_seq_update_in;
isv_status isv_seq_update(isv_handle* h, isv_seq* s, const isv_seq_update_in* in);
/* old_R0 [n][9] = Rs[0] before the solve, pose0 [n][7] = para_Pose[0]; rot_diff_out [n][9] or NULL */
isv_status isv_seq_yaw(isv_handle* h, isv_seq* s, const double* old_R0, const double* pose0, double* rot_diff_out);

typedef struct isv_seq_frame {
  int32_t ex_pose_shared;
  const int64_t* lm_offset;      /* as isv_batch_in                                                 */
  const double* lm_obs;
  int64_t lm_stride;
  const double* pose_fwd;        /* [n][2][7]                                                       */
  const double* ex_pose;
  const double* pose_bwd;        /* [n][2][7]                                                       */
  const double* sb_bwd;          /* [n][2][9]                                                       */
  const double* preint;          /* [n][467]                                                        */
  const double* ts;              /* [n]    Headers[0]   } CombinedFactors members that do not come  */
  const double* Ri;              /* [n][9] Rs[0]        } out of the kernels (:1277-1279); all three */
  const double* ti;              /* [n][3] Ps[0]        } NULL = no pose-graph accumulation          */
} isv_seq_frame;
/* kf_out [n][ISV_ACC_REC] receives the accumulated factor of every sequence whose kf_flag [n] is set
 * (accumFactor->distance > pg_cut_distance; the reference uses 0.1); both may be NULL.             */
isv_status isv_seq_marginalize(isv_handle* h, isv_seq* s, const isv_seq_frame* f, double pg_cut_distance,
                               double* kf_out, int32_t* kf_flag);

/* blocking copies of the whole state to / from HOST memory (checkpoint / resume, tests).  Edge-major:
 * rel [V][n][48] (slot 0 unused), rp [V][n][13] + rp_valid [V][n] (slot = factor index).  `last_*` are
 * the outputs of the most recent isv_seq_marginalize.  NULL members are skipped.                   */
typedef struct isv_seq_host {
  double* rel; double* se3; double* vb; double* rp; int32_t* rp_valid; double* acc; int32_t* pg_count;
  double* last_se3; double* last_pg; double* last_rel; double* last_vb; double* last_rp;
  int32_t* last_rank; int32_t* last_status;
} isv_seq_host;
isv_status isv_seq_export_host(isv_handle* h, isv_seq* s, const isv_seq_host* out);
isv_status isv_seq_import_host(isv_handle* h, isv_seq* s, const isv_seq_host* in);

/* ---- generic marginalization: the engine behind a VINS-Mono style `MarginalizationInfo` ------------
 * (addResidualBlockInfo / preMarginalize / marginalize / getParameterBlocks).  IS-VINS deleted that class
 * (SURVEY.md section 0), so this follows VINS-Mono's published `MarginalizationInfo::marginalize`
 * (vins_estimator/src/factor/marginalization_factor.cpp): A = sum J^T J, b = sum J^T r over every residual
 * block ("ThreadsConstructA"), Schur complement with the eigen-thresholded pseudo-inverse of A_mm,
 * eigen-decomposition of the reduced system, linearized_jacobians = S^1/2 V^T, linearized_residuals =
 * S^-1/2 V^T b.  Parity is unpinned by the reference.
 * Inputs are EVALUATED residual blocks: `values` holds what CostFunction::Evaluate (+ loss correction)
 * wrote -- e.g. the output buffers of isv_eval_problem -- and the tables say where each block lives and
 * which tangent position it maps to.  Tangent layout per problem: [ m_dense | m_diag | n_keep ], pos =
 * their sum; the m_diag marginalized blocks must be scalars that no residual block couples with each
 * other (inverse depths), m_dense <= 32.  All pointers are DEVICE pointers.                          */
typedef struct isv_ne_block {
  int64_t jac_offset;            /* first element of the Jacobian block in `values` (row-major)       */
  int32_t row_stride;            /* doubles between rows (global size: 7 for a pose block)            */
  int32_t local_size;            /* tangent columns used (6 for a pose block), <= 9                   */
  int32_t pos;                   /* position of the block in the tangent vector                       */
  int32_t reserved;
} isv_ne_block;
typedef struct isv_ne_factor {
  int64_t res_offset;            /* first residual in `values`                                        */
  int32_t n_res;                 /* <= 15                                                             */
  int32_t n_blocks;              /* <= 5 (constant parameter blocks are simply not listed)            */
  int32_t first_block;           /* index into the block table                                        */
  int32_t problem;               /* which problem of the batch                                        */
} isv_ne_factor;
typedef struct isv_marg_generic_in {
  int32_t n_problems, pos, m_dense, m_diag;
  int64_t n_factors;
  const isv_ne_factor* factors;
  const isv_ne_block* blocks;
  const double* values;
  double eps;                    /* 1e-8 in VINS-Mono                                                 */
} isv_marg_generic_in;
typedef struct isv_marg_generic_out {
  double* A;                     /* [n_problems][pos][pos] column-major; scratch on return            */
  double* b;                     /* [n_problems][pos]                                                 */
  double* A_red;                 /* [n_problems][n][n]  A_rr - A_rm A_mm^+ A_mr  (n = n_keep)         */
  double* b_red;                 /* [n_problems][n]                                                   */
  double* linearized_jacobians;  /* [n_problems][n][n] column-major, rows in ascending eigenvalue order */
  double* linearized_residuals;  /* [n_problems][n]                                                   */
  int32_t* rank;                 /* [n_problems] eigenvalues of the reduced system > eps              */
  int32_t* status;               /* [n_problems] ISV_W_* (may be NULL)                                */
} isv_marg_generic_out;
/* stage 1 only: zero A, b and scatter-add every residual block (the normal equations of the problem) */
isv_status isv_build_normal_equations(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out);
/* stage 1 + Schur complement + eigen-decomposition */
isv_status isv_marginalize_generic(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out);
/* stage 1 + Schur complement only (A_red, b_red; rank = -1; linearized_* untouched and may be NULL).  With
 * every feature in the diagonal block and m_dense = 0 this is the reduced camera system that ceres'
 * DENSE_SCHUR forms inside problemSolve() (src/estimator.cpp:1124) -- SURVEY.md 8f rank 1.           */
isv_status isv_reduced_system(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out);

/* ---- MarginalizationFactor: the previous round's prior as a residual block ------------------------------
 * VINS-Mono marginalization_factor.cpp `MarginalizationFactor::Evaluate` (IS-VINS deleted the class together
 * with MarginalizationInfo; it is what lets the facade run frame after frame):
 *     dx_b = x_b - x0_b                                   Euclidean blocks
 *     dx_b = [ p - p0 ; +-2 vec(q0^-1 * q) ]              pose blocks (sign: minus when the w of q0^-1 q is < 0)
 *     residual = linearized_residuals + linearized_jacobians * dx
 *     jacobians[b] (n x global_size, row-major) = [ linearized_jacobians[:, idx_b : idx_b + local] | 0 ]
 * All pointers are DEVICE pointers.                                                                       */
typedef struct isv_prior_block {
  int32_t global_size;           /* 7 = pose block (6 tangent columns); anything else is Euclidean           */
  int32_t idx;                   /* first column of the block in the prior (keep_block_idx - m)              */
  int32_t x_offset;              /* first double of the block in `x` and in `x0`                             */
  int32_t pos;                   /* tangent position in the NEW problem (isv_add_marg_prior), -1 = constant  */
} isv_prior_block;
typedef struct isv_marg_prior {
  int32_t n;                     /* rows = columns of linearized_jacobians                                   */
  int32_t n_blocks;
  const isv_prior_block* blocks;
  const double* linearized_jacobians;   /* [n][n] column-major (as isv_marg_generic_out writes it)          */
  const double* linearized_residuals;   /* [n]                                                               */
  const double* x0;              /* keep_block_data: the kept blocks at the linearization point              */
  const double* x;               /* the same blocks now                                                      */
} isv_marg_prior;
/* residuals [n]; jacobians (may be NULL): the blocks concatenated in table order, block b = n x global_size
 * row-major.  ISV_W_BAD_INDEX in *status (may be NULL) when a block reaches outside the prior.             */
isv_status isv_eval_marg_prior(isv_handle* h, const isv_marg_prior* prior, double* residuals, double* jacobians,
                               int32_t* status);
/* A += J^T J, b += J^T residuals of the prior, scattered by blocks[].pos into the normal equations of problem
 * `problem` (out->A, out->b as isv_build_normal_equations left them; problem = -1: the same prior into every
 * problem of the batch); residuals from isv_eval_marg_prior.                                                */
isv_status isv_add_marg_prior(isv_handle* h, const isv_marg_prior* prior, const double* residuals,
                              const isv_marg_generic_in* in, const isv_marg_generic_out* out, int32_t problem);
/* stage 2 alone on normal equations that are already built (isv_build_normal_equations [+ isv_add_marg_prior]):
 * Schur complement (+ eigen-decomposition unless schur_only).  build -> add prior -> this == what
 * isv_marginalize_generic does in one call when there is no prior.                                          */
isv_status isv_schur_eig(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out,
                         int32_t schur_only);

/* The whole MarginalizationInfo::preMarginalize + marginalize step from HOST memory, blocking: H2D of the
 * parameter blocks and factor lists, `Evaluate` of every residual block on the GPU (isv_eval_*), the block
 * tables built from the per-family tangent positions, isv_marginalize_generic (or isv_reduced_system when
 * schur_only != 0), D2H of the results.  This is what the C++ `MarginalizationInfo` of
 * is_vins_b200/host/isv_marginalization_info.hpp calls.  All pointers are HOST pointers; the factor
 * structs are the ones of the isv_eval_* family.  pos_* give the tangent position of every parameter block
 * in the layout [m_dense | m_diag | keep] or -1 for a block ceres holds constant / that no factor uses.   */
typedef struct isv_marg_host_in {
  isv_param_blocks pb;
  isv_proj_factors proj;
  isv_imu_factors imu;
  isv_small_factors small_factors;
  const int32_t* pos_pose;
  const int32_t* pos_speed_bias;
  const int32_t* pos_ex_pose;
  const int32_t* pos_feature;
  int32_t pos, m_dense, m_diag, schur_only;
  double eps;
  /* ProjectionTdFactor problems (proj.td_obs != NULL; proj.td / proj.td_idx are HOST arrays here): tangent
   * position of every para_Td block, -1 = held constant.  The time offset is never a diagonal marginalized
   * scalar (it couples with every visual factor).                                                          */
  const int32_t* pos_td;         /* [proj.n_td]                                                          */
  /* MarginalizationFactor: the previous round's prior as one more residual block (at most one, as in
   * VINS-Mono's Estimator::optimization), or NULL.  HOST pointers inside; blocks[].pos = tangent position of
   * each kept block in THIS problem (-1 = constant), x = the blocks' current values.                      */
  const isv_marg_prior* prior;
} isv_marg_host_in;
typedef struct isv_marg_host_out {
  double* A_red;                 /* [n][n] column-major, n = pos - m_dense - m_diag                    */
  double* b_red;                 /* [n]                                                                */
  double* linearized_jacobians;  /* [n][n] column-major (untouched when schur_only)                    */
  double* linearized_residuals;  /* [n]                                                                */
  int32_t rank;
  int32_t status;
} isv_marg_host_out;
isv_status isv_marginalize_host(isv_handle* h, const isv_marg_host_in* in, isv_marg_host_out* out);

/* ---- PoseLocalParameterization  src/factor/pose_local_parameterization.cpp:3-27 ------------------------------
 * Plus: x_plus_delta = [p + dp ; normalize(q * deltaQ(dtheta))] for n pose blocks (x [n][7], delta [n][6],
 * x_plus_delta [n][7]; DEVICE pointers; in place allowed) -- what ceres applies to para_Pose / para_Ex_Pose after a
 * step, batched over every pose block of every window.  ComputeJacobian is the constant [I6 ; 0] (7 x 6 row-major):
 * written to `jacobian` [42] (HOST pointer).                                                                    */
isv_status isv_pose_plus_batch(isv_handle* h, int64_t n, const double* x, const double* delta, double* x_plus_delta);
void isv_pose_plus_jacobian(double* jacobian);

/* ---- unit-test hook: the PSD eigensolver that replaces SelfAdjointEigenSolver on this path -----
 * (src/estimator.cpp:920,1311,1479).  nb symmetric n x n matrices A (column-major, host) ->
 * G [nb][n][n] row-major factor rows with  A ~= sum_k g_k g_k^T, g_k mutually orthogonal;
 * lam [nb][n] = |g_k|^2 (eigenvalues, zero-padded); info [nb][2] = {rows, sweeps}.  n <= 63.      */
isv_status isv_test_psd_eig(isv_handle* h, int nb, int n, const double* A, double* G, double* lam,
                            int32_t* info);

/* ---- unit-test hook: the lean reciprocal square root / reciprocal the hot chains use instead of rsqrt() and 1.0 / x
 * (hardware seed + two Newton steps, is_vins_b200/csrc/isv_device_math.cuh): x [n] (host) -> 1 / sqrt(x), 1 / x.       */
isv_status isv_test_fast_special(isv_handle* h, int n, const double* x, double* rsqrt_out, double* rcp_out);

/* ---- measurement hook: `iters` back-to-back isv_marg_event calls of the same event, timed one by one on the host with
 * CLOCK_MONOTONIC from inside the library (i.e. the latency a C/C++ estimator sees, without a binding's marshalling).
 * us [iters] = microseconds of each call; the last call's results are left in fwd_out / bwd_out.                    */
isv_status isv_test_event_latency(isv_handle* h, const isv_fwd_in* fwd_in, const isv_bwd_in* bwd_in, isv_fwd_out* fwd_out,
                                  isv_bwd_out* bwd_out, int iters, double* us);

/* ---- profiling hook: the fused single-event kernel on a DEVICE batch with the SM cycle counter sampled at the phase
 * boundaries of the eight warps of window 0.  stamps_dev: device int64 [8 warps][8]: [0] entry, [1] after the start barrier,
 * then per role -- warp 0: [2] end of MargBackward; warps 1-7: [2] factor-Jacobian task done, [3] landmark share done;
 * warp 7: [4] forward barrier passed, [5] tail done; [7] = after the final barrier (all warps).                        */
isv_status isv_test_fused_stamps(isv_handle* h, const isv_batch_in* in, const isv_batch_out* out, int64_t* stamps_dev);

/* ---- unit-test hook: the symmetric eigensolver of the generic engine's reduced system (VINS-Mono
 * `SelfAdjointEigenSolver<MatrixXd> saes2(A)`; Householder tridiagonalization + implicit QL with a rotation log,
 * is_vins_b200/csrc/isv_sym_eig.cuh).  nb symmetric n x n matrices A (column-major, host) -> lam [nb][n] ascending,
 * V [nb][n][n] column-major with column r the eigenvector of lam[r]; info [nb][2] = {QL failure flags, rotations
 * logged}.  n <= 1024.                                                                              */
isv_status isv_test_sym_eig(isv_handle* h, int nb, int n, const double* A, double* lam, double* V, int32_t* info);

#ifdef __cplusplus
}
#endif
#endif /* ISV_CAPI_H */
