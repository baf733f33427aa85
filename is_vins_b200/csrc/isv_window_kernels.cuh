// The two window kernels of the hot path: one WARP per sliding window, every stage fused, so that
// only the window's inputs and its recovered factors touch HBM (SURVEY.md 7.2 "KF").
//
//   marg_forward_accum_kernel + marg_forward_tail_kernel
//                         Estimator::MargForward    /root/reference/src/estimator.cpp:1149-1352
//   marg_backward_kernel  Estimator::MargBackward   /root/reference/src/estimator.cpp:1354-1539
//
// MargForward, structured form.  The reference builds the dense (12+L)^2 `Lamda` with the block
// loop :1168-1202 and inverts the (L+6)^2 block with FullPivLU (:1286).  The landmark block of
// Lamda is diagonal (every ProjectionFactor touches exactly one inverse depth), so eliminating the
// landmarks first is algebraically identical and O(L):
//     s*J_k = [a_k; c_k] (2x12, columns ordered [T1|T0] as OrderMap :1153-1158), s*jl_k (2x1)
//     u_k = s*jl_k/|s*jl_k| , v_k _|_ u_k ;  e_k = (sJ_k)^T u_k , w_k = (sJ_k)^T v_k
//     Lamda[0:12,0:12]            = sum_k e_k e_k^T + w_k w_k^T   (+ prior + rel-pose blocks)
//     Lamda[0:12,0:12] - B D^-1 B^T = sum_k w_k w_k^T             (Schur over the landmarks)
// Both sums are SYRKs  X X^T  with X 12 x L of rank <= 6: they are accumulated as 6x6 Gram matrices
// in registers (see "6-vector form" below).  Everything after that is 12x12 / 6x6 algebra.
#pragma once
#include <type_traits>

#include "isv_device_math.cuh"
#include "isv_factors.cuh"
#include "isv_small_qr.cuh"
#include "isv_warp_linalg.cuh"

#include "../../include/isv_capi.h"

namespace isv {

#ifndef ISV_WARPS_PER_CTA
#define ISV_WARPS_PER_CTA 4
#endif
constexpr int kWarpsPerCta = ISV_WARPS_PER_CTA;
constexpr int kThreads = 32 * kWarpsPerCta;
#ifndef ISV_FWD_MINB
#define ISV_FWD_MINB 2
#endif
#ifndef ISV_FWD_TAIL_MINB
#define ISV_FWD_TAIL_MINB 4
#endif
#ifndef ISV_BWD_MINB
#define ISV_BWD_MINB 4
#endif

// ---- forward: shared-memory map (doubles, per warp) -----------------------------------------
constexpr int kFwdConst = 72;                // Psi (12 x 6)
constexpr int kFwdWork = 832;                // tail matrices
constexpr int kFwdSmemPerWarp = kFwdConst + kFwdWork;
// ---- backward ---------------------------------------------------------------------------------
constexpr int kBwdSmemPerWarp = 272 + 24 + 48 + 316 + 472 + 48;   // Lc | hv | pivot rows | Gs | T | sc (all even: 16-byte aligned)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int nonunit(const double* ps) {
  double n2 = ps[3] * ps[3] + ps[4] * ps[4] + ps[5] * ps[5] + ps[6] * ps[6];
  return fabs(n2 - 1.0) > 1e-9;
}

// out (n x n col-major, global) = upper-triangular Eigen `LLT(M).matrixL().transpose()`;
// M (smem, ld) is destroyed.  Returns 1 when M is not SPD.
__device__ __forceinline__ int chol_store_upper(double* M, int ld, int n, double* out, int lane, int& nonfinite) {
  int bad = w_chol_lower(M, ld, n, lane);
  for (int idx = lane; idx < n * n; idx += 32) {
    int i = idx % n, j = idx / n;
    double v = (i <= j) ? M[j + i * ld] : 0.0;
    if (!isfinite(v)) nonfinite = 1;
    out[idx] = v;
  }
  __syncwarp();
  return bad;
}

// =================================================================================================
// Factor-Jacobian pre-kernel.  Everything in Marg* that is ONE short scalar chain per window -- the
// tangent twins of the prior / recovered factors and of the IMU factor (SO(3) log / Jr^-1 with their
// transcendental calls), the records that are pure functions of the poses -- is evaluated here with
// one THREAD per window instead of one lane of a warp per window, and handed to the window kernels
// through a [n][kFJ] scratch.  Inside the warp-per-window kernels the same work kept 31 lanes idle for
// ~1.5-1.8 k instructions of a latency-bound chain.  Measured alternatives that were slower: the IMU factor split 9 ways over threads
// (redundant strided header loads), a coalesced warp-per-window IMU kernel (45 us), __launch_bounds__
// for more resident CTAs (spills).
//   tasks 0-3 (forward):  Wst (12x12) = [ sp [0 | Jp] ; sr [Jj | Ji] ]  (:1203-1238),
//       G (6x12) = [Ji | Jj] of the pose-graph RelativePoseFactor (:1244-1255), Jr6 of the new
//       SE3PriorFactor (:1291-1297); writes pg_out[0:12), [84], [85:89) and se3_out[0:12)
//   tasks 4-6 (backward): IMU Jacobian 15x30 row-major in OrderMap column order (:1382-1412),
//       [Ji | Jj] of the new RelativePoseFactor (:1424-1433), J of the RollPitchFactor (:1445-1447);
//       writes rel_out[0:12), rp_out[0:9), vb_out[0:9)
// =================================================================================================
constexpr int kFJ_W = 0, kFJ_G = 144, kFJ_JR6 = 216, kFJ_IMU = 252, kFJ_REL = 702, kFJ_RP = 774, kFJ = 786;

constexpr int kJacTasks = 7;   // blockIdx.y: one short chain per (window, factor)

// One (window, task) chain, executed by one thread.  F = the window's factor-Jacobian record (kFJ doubles; global scratch in
// the batch kernels, shared memory in the fused single-event kernel).  wstatus: where status bits are OR-ed (may be null).
__device__ __forceinline__ void factor_jac_task(const isv_batch_in& in, const isv_batch_out& out, double* F, const DevCfg& cfg,
                                                int win, int task, int32_t* wstatus) {
  if (task < 4) {
    const double* pose0 = in.pose_fwd + (size_t)win * 14;
    const double* pose1 = pose0 + 7;
    const double* pse3 = in.prior_se3 + (size_t)win * ISV_SE3_REC;
    const double* prel = in.prior_rel + (size_t)win * ISV_REL_REC;
    double* o_se3 = out.se3_out + (size_t)win * ISV_SE3_REC;
    double* o_pg = out.pg_out + (size_t)win * ISV_PG_REC;
    double* W = F + kFJ_W;
    if (task == 0) {
      // vioRelativePoseEdges[1]: rows 6-11 = sr * [Jj | Ji]   (OrderMap: T1@0, T0@6)
      // (a thread per window reads its record with a 384-byte stride: load every value exactly once)
      double Ja[36], Jb[36], dt[3], dR[9], sr[36];
#pragma unroll
      for (int i = 0; i < 36; ++i) sr[i] = prel[12 + i];
      for (int i = 0; i < 3; ++i) dt[i] = prel[i];
      load_mat3_colmajor(prel + 3, dR);
      relpose_jacobians(pose0, pose1, dt, dR, Ja, Jb, nullptr);
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        const double* Jc = (c < 6) ? (Jb + 6 * c) : (Ja + 6 * (c - 6));
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          double acc = 0.0;
#pragma unroll
          for (int l = 0; l < 6; ++l) acc = fma(sr[r + 6 * l], Jc[l], acc);
          W[6 + r + 12 * c] = acc;
        }
      }
    } else if (task == 1) {
      // vioPosePriorEdge: rows 0-5 = sp * [0 | Jp]
      double Ja[36], tt[3], Rp[9], sp[36];
#pragma unroll
      for (int i = 0; i < 36; ++i) sp[i] = pse3[12 + i];
      for (int i = 0; i < 3; ++i) tt[i] = pse3[i];
      load_mat3_colmajor(pse3 + 3, Rp);
      se3prior_jacobian(pose0, tt, Rp, Ja, nullptr);
#pragma unroll
      for (int c = 0; c < 12; ++c)
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          double acc = 0.0;
          if (c >= 6) {
#pragma unroll
            for (int l = 0; l < 6; ++l) acc = fma(sp[r + 6 * l], Ja[l + 6 * (c - 6)], acc);
          }
          W[r + 12 * c] = acc;
        }
    } else if (task == 2) {
      // pose-graph RelativePoseFactor(T0 -> T1) at the current estimate
      double dt[3], dR[9];
      Quat Qi = quat_from_pose(pose0), Qj = quat_from_pose(pose1);
      double dd[3] = {pose1[0] - pose0[0], pose1[1] - pose0[1], pose1[2] - pose0[2]};
      qrot(qinv(Qi), dd, dt);
      q2R(qmul(qinv(Qi), Qj), dR);
      for (int i = 0; i < 3; ++i) o_pg[i] = dt[i];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o_pg[3 + r + 3 * c] = dR[3 * r + c];
      o_pg[84] = sqrt(dt[0] * dt[0] + dt[1] * dt[1] + dt[2] * dt[2]);  // distance = delta_t.norm()
      relpose_jacobians(pose0, pose1, dt, dR, F + kFJ_G, F + kFJ_G + 36, nullptr);
    } else {
      // new SE3PriorFactor(P1, Q1)
      double tt[3], Rp[9];
      q2R(quat_from_pose(pose1), Rp);
      for (int i = 0; i < 3; ++i) { tt[i] = pose1[i]; o_se3[i] = tt[i]; }
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o_se3[3 + r + 3 * c] = Rp[3 * r + c];
      se3prior_jacobian(pose1, tt, Rp, F + kFJ_JR6, nullptr);
      // covAbs = (s^T s)^-1 of vioRollPitchEdges[0] when its index is 0 (:1265-1271)
      double cA[4] = {0, 0, 0, 0};
      if (in.prior_rp) {
        const double* rp = in.prior_rp + (size_t)win * ISV_RP_IN_REC;
        if (rp[0] != 0.0) {
          double a = rp[1], b = rp[2], c = rp[3], d = rp[4];  // col-major s = [a c; b d]
          double m00 = a * a + b * b, m01 = a * c + b * d, m11 = c * c + d * d;
          double det = m00 * m11 - m01 * m01;
          cA[0] = m11 / det; cA[1] = -m01 / det; cA[2] = -m01 / det; cA[3] = m00 / det;
        }
      }
      for (int i = 0; i < 4; ++i) o_pg[85 + i] = cA[i];
    }
  } else {
    const double* pose_i = in.pose_bwd + (size_t)win * 14;
    const double* pose_j = pose_i + 7;
    const double* sb_i = in.sb_bwd + (size_t)win * 18;
    const double* sb_j = sb_i + 9;
    if (task == 4) {
      // IMUFactor::Evaluate, tangent twin (imu_factor.h:161-265); OrderMap (:1358-1366):
      // T_V@0, VB_V@6, T_{V-1}@15, VB_{V-1}@21.  The record was zero-filled by the launcher.
      const double* pre = in.preint + (size_t)win * ISV_PREINT_REC;
      if ((nonunit(pose_i) || nonunit(pose_j)) && wstatus) atomicOr(wstatus, ISV_W_NONUNIT_QUAT);
      imu_jacobians(pose_i, sb_i, pose_j, sb_j, pre, cfg.g, F + kFJ_IMU, 1, 15, 21, 0, 6, nullptr, 0, 1, 30);
    } else if (task == 5) {
      double* o_rel = out.rel_out + (size_t)win * ISV_REL_REC;
      Quat Qi = quat_from_pose(pose_i), Qj = quat_from_pose(pose_j);
      double dd[3] = {pose_j[0] - pose_i[0], pose_j[1] - pose_i[1], pose_j[2] - pose_i[2]};
      double tij[3], Rij[9];
      qrot(qinv(Qi), dd, tij);
      q2R(qmul(qinv(Qi), Qj), Rij);
      relpose_jacobians(pose_i, pose_j, tij, Rij, F + kFJ_REL, F + kFJ_REL + 36, nullptr);
      for (int i = 0; i < 3; ++i) o_rel[i] = tij[i];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o_rel[3 + r + 3 * c] = Rij[3 * r + c];
    } else {
      // RollPitchFactor(Qw) with Qw = Q_{V-1}: member R = Qw.toRotationMatrix()
      double* o_vb = out.vb_out + (size_t)win * ISV_VB_REC;
      double* o_rp = out.rp_out + (size_t)win * ISV_RP_REC;
      double Rm[9];
      q2R(quat_from_pose(pose_i), Rm);
      rollpitch_jacobian(pose_i, Rm, F + kFJ_RP, nullptr);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o_rp[r + 3 * c] = Rm[3 * r + c];
      for (int i = 0; i < 9; ++i) o_vb[i] = sb_j[i];   // Linear9Factor(vb): VB = para_SpeedBias[V]
    }
  }
}

// Zero-fill helpers of the batch launcher.  They replace cudaMemsetAsync / cudaMemset2DAsync, which the driver may hand to a
// copy engine: inside the chunked host pipeline (isv_marg_window_batch_host) such a memset queued behind the next chunk's
// multi-megabyte H2D copies and held the kernel chain up for up to 0.4 ms (profiles/r02t_host_pipeline_trace.txt).
__global__ void zero_rows_kernel(double* __restrict__ base, int n_rows, int row_len, int row_stride) {
  const long long total = (long long)n_rows * row_len;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row_len;
    base[r * row_stride + (i - r * row_len)] = 0.0;
  }
}
__global__ void zero_i32_kernel(int32_t* __restrict__ p, long long n, int32_t* __restrict__ q = nullptr, int nq = 0) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = 0;
  if (blockIdx.x == 0 && (int)threadIdx.x < nq) q[threadIdx.x] = 0;
}

// ABI 4 (host path): records without their structural zeros.  A record is up to four segments, each either copied or an
// N x N column-major block whose upper triangle travels (column by column: (0,0), (0,1), (1,1), (0,2), ...).
// PACK = false: packed -> full (strict lower triangle = 0); PACK = true: full -> packed.  One thread per full element.
struct TriLayout { int nseg; int kind[4]; int len[4]; };   // kind 0: copy `len` doubles; kind 1: upper triangle of a len x len block
__host__ __device__ inline int tri_full_len(const TriLayout& L) { int t = 0; for (int s = 0; s < L.nseg; ++s) t += L.kind[s] ? L.len[s] * L.len[s] : L.len[s]; return t; }
__host__ __device__ inline int tri_packed_len(const TriLayout& L) { int t = 0; for (int s = 0; s < L.nseg; ++s) t += L.kind[s] ? L.len[s] * (L.len[s] + 1) / 2 : L.len[s]; return t; }
template <bool PACK>
__global__ void tri_records_kernel(long long n, const double* __restrict__ src, double* __restrict__ dst, TriLayout lay) {
  const int fl = tri_full_len(lay), pl = tri_packed_len(lay);
  const long long total = n * fl;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long w = idx / fl;
    int e = (int)(idx - w * fl), fo = 0, po = 0, pe = -1;
    for (int s = 0; s < lay.nseg; ++s) {
      const int N = lay.len[s];
      const int fs = lay.kind[s] ? N * N : N, ps = lay.kind[s] ? N * (N + 1) / 2 : N;
      if (e < fo + fs) {
        const int r = e - fo;
        if (!lay.kind[s]) pe = po + r;
        else { const int i = r % N, j = r / N; pe = (i <= j) ? po + j * (j + 1) / 2 + i : -1; }
        break;
      }
      fo += fs; po += ps;
    }
    if (PACK) { if (pe >= 0) dst[w * pl + pe] = src[idx]; }
    else dst[idx] = pe >= 0 ? src[w * pl + pe] : 0.0;
  }
}

__global__ void __launch_bounds__(128)
marg_factor_jac_kernel(isv_batch_in in, isv_batch_out out, double* __restrict__ fj, DevCfg cfg, int task0) {
  const int win = blockIdx.x * blockDim.x + threadIdx.x;
  if (win >= in.n_windows) return;
  const int task = task0 + blockIdx.y;   // forward launch: task0 = 0, gridDim.y = 4; backward: task0 = 4, gridDim.y = 3
  factor_jac_task(in, out, fj + (size_t)win * kFJ, cfg, win, task, out.status ? out.status + win : nullptr);
}

// =================================================================================================
// MargForward
//
// 6-vector form.  A ProjectionFactor's residual depends on the two poses only through the relative
// pose T0^-1 T1, so every weighted Jacobian row lives in a 6-dimensional space.  With
//   D = R0^T R1 ,  F = R1^T R0 ric ,  f = R1^T R0 tic ,  d = R1^T (P0 - P1) ,  tp = ric^T (f + d - tic)
// and, per landmark (p = pts_i, lam = inv_dep):
//   w = F p ;  pi^ = w + lam f ;  q~ = ric^T w ;  c~ = q~ + lam tp  (= lam * pts_camera_j) ;
//   rho = 1 / c~_z ;  (xb, yb) = c~_xy rho ;  Pb = [1 0 -xb ; 0 1 -yb]
//   u = normalize(s Pb q~)  (direction of s * jacobian_feature, :176-178) , v _|_ u
//   alpha = ric Pb^T (rho s^T u) ,  beta = ric Pb^T (rho s^T v)
//   y_e = [ pi^ x alpha ; lam alpha ] ,  y_w = [ pi^ x beta ; lam beta ]                 (6-vectors)
// the 12-vectors of the 9/12-dimensional forms are  e = Psi y_e , w = Psi y_w  with the per-window
//   Psi (12x6) = [ 0 -R1 ; -I -skew(d) ; 0 R1 ; D 0 ]     (rows: OrderMap T1@0 [pos, rot], T0@6)
// so  Lamda[0:12,0:12] = Psi (sum y_e y_e^T + y_w y_w^T) Psi^T  and the landmark Schur complement
// is  Psi (sum y_w y_w^T) Psi^T.  The two 6x6 Gram matrices (21 + 21 unique entries) are accumulated
// in registers, one landmark per lane: ~100 DFMA for the chain + 42 for the SYRKs per landmark and
// no shared-memory staging.  (History: the 9-vector version ran the SYRKs as DMMA.8x8x4 tiles; on
// B200 DMMA and DFMA share one pipe at the same FLOP rate (tools/fp64_peak.cu), an 8x8 tile wastes
// 64/21 of the work on a 6x6 symmetric Gram, and the kernel is FP64-pipe bound -- DFMA is 3.4x faster
// here.  DMMA stays where the tile is full: the dense Schur product of the generic path.)
// =================================================================================================
// ---- kernel 1: landmark phase.  One warp per window, one landmark per lane and iteration;
// writes the 21 + 21 lower-triangle entries of the two Gram matrices to gram[win][42].
constexpr int kAccLd = 33;                         // reduction staging: [21][33] doubles per warp
constexpr int kAccSmemPerWarp = 34 + 21 * kAccLd + 1;  // constants F f tp ric M + staging (even: 16-byte aligned per warp)

// ZONE = ISV_IN_PTS_I_Z_ONE: the caller guarantees pts_i.z == 1 (src/System.cpp:346); component 2 is then not read
// Warps per CTA of the landmark kernel.  A warp owns a whole window, so a CTA lives as long as its LONGEST window: with
// ragged landmark counts (L ~ U{0.75 .. 1.25} mean) four windows per CTA idle ~15 % of their slots waiting for the
// largest one.  One warp per CTA removes that, and 224 registers x 32 threads also packs 9 instead of 8 warps per SM.
#ifndef ISV_ACC_WARPS
#define ISV_ACC_WARPS 1
#endif
#ifndef ISV_ACC_MQ
#define ISV_ACC_MQ 0   // q~ = (ric^T F) p straight from the observation: shorter chain, but measured slower (register pressure)
#endif
constexpr int kAccWarps = ISV_ACC_WARPS;
#ifndef ISV_ACC_MINB
#define ISV_ACC_MINB (ISV_FWD_MINB * kWarpsPerCta / kAccWarps)
#endif

#ifdef ISV_ACC_MAXNREG
#define ISV_ACC_BOUNDS __maxnreg__(ISV_ACC_MAXNREG)
#else
#define ISV_ACC_BOUNDS __launch_bounds__(32 * kAccWarps, ISV_ACC_MINB)
#endif
// ISO = ProjectionFactor::sqrt_info is a multiple of the identity (it always is in the reference: FOCAL_LENGTH / 1.5 * I,
//       src/estimator.cpp:35): the weighting commutes with the direction and drops out of the per-landmark chain.
// Body of the landmark phase for one warp: landmarks [lm0, lm0 + L) of window `win` -> the 42 Gram-triangle entries at g.
// The batch kernel gives a warp its whole window; the fused single-event kernel splits one window over several warps.
// K = kAccSmemPerWarp doubles of shared memory owned by the calling warp; wstatus = where status bits are OR-ed (or null).
// XYF = pts_i.x / pts_i.y come as the FP32 values the feature tracker produced (isv_batch_in::lm_xy_f32), widened on load
template <bool ZONE, bool ISO, bool XYF = false>
__device__ __forceinline__ void forward_accum_body(const isv_batch_in& in, const DevCfg& cfg, const int win, const int lane,
                                                   double* K, const long long lm0, const int L, double* __restrict__ g,
                                                   int32_t* wstatus, const int lam_comp = 5) {
  double* R = K + 34;                         // reduction staging;  K: [0]F [9]f [12]tp [15]ric [24]M = ric^T F
  int status = 0;

  const double* pose0 = in.pose_fwd + (size_t)win * 14;
  const double* pose1 = pose0 + 7;
  const double* ex = in.ex_pose + (in.ex_pose_shared ? 0 : (size_t)win * 7);
  if (lane == 0) {
    Quat Qi = quat_from_pose(pose0), Qj = quat_from_pose(pose1), qic = quat_from_pose(ex);
    if (nonunit(pose0) || nonunit(pose1) || nonunit(ex)) status |= ISV_W_NONUNIT_QUAT;
    double ric[9], R0[9], R1[9], D[9], F[9], fv[3], dv[3], tp[3], t3[3];
    q2R(qic, ric);
    q2R(Qi, R0);
    q2R(Qj, R1);
    mat3_tmul(R0, R1, D);          // D = R0^T R1
    mat3_tmul(D, ric, F);          // F = D^T ric = R1^T R0 ric
    mat3_tvec(D, ex, fv);          // f = D^T tic
    for (int i = 0; i < 3; ++i) t3[i] = pose0[i] - pose1[i];
    mat3_tvec(R1, t3, dv);         // d = R1^T (P0 - P1)
    for (int i = 0; i < 3; ++i) t3[i] = fv[i] + dv[i] - ex[i];
    mat3_tvec(ric, t3, tp);
    for (int i = 0; i < 9; ++i) { K[i] = F[i]; K[15 + i] = ric[i]; }
    for (int i = 0; i < 3; ++i) { K[9 + i] = fv[i]; K[12 + i] = tp[i]; }
    double Mq[9];
    mat3_tmul(ric, F, Mq);         // M = ric^T F: q~ = M p straight from the observation (not through w = F p)
    for (int i = 0; i < 9; ++i) K[24 + i] = Mq[i];
  }
  __syncwarp();

  const long long st = in.lm_stride;
  const double s00 = cfg.ps[0], s10 = cfg.ps[1], s01 = cfg.ps[2], s11 = cfg.ps[3];
  // ric is used three times per landmark: keep it in registers; F, f, tp are broadcast LDS
  double ric[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) ric[i] = K[15 + i];

  // 21 + 21 accumulators: lower triangles of sum y_e y_e^T and sum y_w y_w^T
  double ae[21], aw[21];
#pragma unroll
  for (int i = 0; i < 21; ++i) { ae[i] = 0.0; aw[i] = 0.0; }

  // one landmark -> its two 6-vectors (see the derivation above); returns 1 for a degenerate observation
  auto chain = [&](double px, double py, double pz, double lam, double* __restrict__ ye, double* __restrict__ yw) -> int {
    double w[3], ph[3], qt[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      w[r] = ZONE ? fma(K[3 * r], px, fma(K[3 * r + 1], py, K[3 * r + 2])) : K[3 * r] * px + K[3 * r + 1] * py + K[3 * r + 2] * pz;
      ph[r] = fma(lam, K[9 + r], w[r]);
#if ISV_ACC_MQ
      qt[r] = ZONE ? fma(K[24 + 3 * r], px, fma(K[25 + 3 * r], py, K[26 + 3 * r])) : K[24 + 3 * r] * px + K[25 + 3 * r] * py + K[26 + 3 * r] * pz;
#endif
    }
#if !ISV_ACC_MQ
#pragma unroll
    for (int c = 0; c < 3; ++c) qt[c] = ric[c] * w[0] + ric[3 + c] * w[1] + ric[6 + c] * w[2];
#endif
    const double c0 = fma(lam, K[12], qt[0]), c1 = fma(lam, K[13], qt[1]), c2 = fma(lam, K[14], qt[2]);
    // direction of s * jacobian_feature:  Pb q~ = [q0 - xb q2, q1 - yb q2] = lam / c2 * [q0 tp2 - q2 tp0,
    // q1 tp2 - q2 tp1]  (c~ = q~ + lam tp).  Only the direction matters (u enters as u u^T, v as v v^T), so it
    // is taken from the cancellation-free right-hand side.
    const double nt0 = fma(qt[0], K[14], -qt[2] * K[12]), nt1 = fma(qt[1], K[14], -qt[2] * K[13]);
    const double g0 = ISO ? nt0 : s00 * nt0 + s01 * nt1, g1 = ISO ? nt1 : s10 * nt0 + s11 * nt1;   // ISO: s = s00 I, direction unchanged
    const double n2 = g0 * g0 + g1 * g1;
    // ONE reciprocal square root gives both 1 / (|g| c2) and rho = 1 / c2 (the FP64 division and rsqrt are the two
    // longest instruction sequences of the chain):  r = rsqrt(n2 c2^2) ;  1 / (|g| c2) = sign(c2) r ;  rho = r^2 n2 c2
    const double r = rsqrt(n2 * (c2 * c2));   // (the lean fast_rsqrt costs this kernel registers: 255 + a spill, +1.7 % time)
    const bool ok = n2 > 0.0;
    double rho = (r * r) * (n2 * c2);
    double in_ = copysign(r, c2);
    if (__builtin_expect(!ok, 0)) { rho = 1.0 / c2; in_ = 0.0; }      // degenerate observation (flagged below)
    const double xb = c0 * rho, yb = c1 * rho;
    if (ISO) in_ *= s00;
    const double u0 = ok ? g0 * in_ : (ISO ? rho * s00 : rho), u1 = ok ? g1 * in_ : 0.0;        // (u, v) pre-scaled by rho [s00]
    // ISO: (u, v) additionally pre-scaled by s00 below, so rho s^T u = u and rho s^T v = (-u1, u0)
    const double a0 = ISO ? u0 : u0 * s00 + u1 * s10, a1 = ISO ? u1 : u0 * s01 + u1 * s11;        // rho s^T u
    const double b0 = ISO ? -u1 : -u1 * s00 + u0 * s10, b1 = ISO ? u0 : -u1 * s01 + u0 * s11;     // rho s^T v
    double al[3], be[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double n0 = fma(-xb, ric[3 * r + 2], ric[3 * r]), n1 = fma(-yb, ric[3 * r + 2], ric[3 * r + 1]);
      al[r] = a0 * n0 + a1 * n1;
      be[r] = b0 * n0 + b1 * n1;
    }
    ye[0] = ph[1] * al[2] - ph[2] * al[1];  ye[1] = ph[2] * al[0] - ph[0] * al[2];  ye[2] = ph[0] * al[1] - ph[1] * al[0];
    yw[0] = ph[1] * be[2] - ph[2] * be[1];  yw[1] = ph[2] * be[0] - ph[0] * be[2];  yw[2] = ph[0] * be[1] - ph[1] * be[0];
#pragma unroll
    for (int r = 0; r < 3; ++r) { ye[3 + r] = lam * al[r]; yw[3 + r] = lam * be[r]; }
    return ok ? 0 : 1;
  };
  auto syrk = [&](const double* __restrict__ ye, const double* __restrict__ yw) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        ae[t] = fma(ye[i], ye[j], ae[t]);
        aw[t] = fma(yw[i], yw[j], aw[t]);
        ++t;
      }
  };
  // Two landmarks per lane and step (k and k + 32): their projection chains are independent, which doubles the
  // instruction-level parallelism of the latency-bound part.  The observations travel through a ring of four register
  // entries (entry e = landmarks base + 32 e + lane); an entry is refilled -- for the landmark 128 further on, i.e. two
  // steps ahead -- as soon as its values have been copied out.  The loop is unrolled over the ring, so ring indices are
  // compile-time (no register shuffling) and the loads use immediate offsets from four running pointers.
  using xy_t = typename std::conditional<XYF, float, double>::type;
  const xy_t* __restrict__ qx = (XYF ? reinterpret_cast<const xy_t*>(in.lm_xy_f32) : reinterpret_cast<const xy_t*>(in.lm_obs)) + lm0 + lane;
  const xy_t* __restrict__ qy = qx + st;
  const double* __restrict__ qz = in.lm_obs + lm0 + lane + 2 * st;
  const double* __restrict__ ql = in.lm_obs + lm0 + lane + lam_comp * st;   // 5 in the ABI's layout; the packed event block: 2 / 3
  xy_t rx[4], ry[4];   // XYF: the ring holds the floats (8 registers less), widened when a step consumes them
  double rz[4], rl[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    rx[e] = 0; ry[e] = 0; rz[e] = 1.0; rl[e] = 1.0;
    if (32 * e + lane < L) { rx[e] = qx[32 * e]; ry[e] = qy[32 * e]; if (!ZONE) rz[e] = qz[32 * e]; rl[e] = ql[32 * e]; }
  }
  // Main loop: whole super-steps of 128 landmarks.  Every lane has work, so the body carries no predicate and no branch:
  // one basic block of four chains and four SYRKs that the scheduler is free to interleave (the SYRK of one step hides
  // the dependent chain of the next).  Only the refills are predicated (the prefetch runs past the window's end).
  int base = 0;
#pragma unroll 1
  for (; base + 128 <= L; base += 128) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double pxa = rx[2 * h], pya = ry[2 * h], pza = rz[2 * h], la = rl[2 * h];
      const double pxb = rx[2 * h + 1], pyb = ry[2 * h + 1], pzb = rz[2 * h + 1], lb = rl[2 * h + 1];
#pragma unroll
      for (int e = 2 * h; e < 2 * h + 2; ++e)
        if (base + 128 + 32 * e + lane < L) {
          rx[e] = qx[128 + 32 * e]; ry[e] = qy[128 + 32 * e]; if (!ZONE) rz[e] = qz[128 + 32 * e]; rl[e] = ql[128 + 32 * e];
        }
      double yea[6], ywa[6], yeb[6], ywb[6];
      int bad = chain(pxa, pya, pza, la, yea, ywa);
      bad |= chain(pxb, pyb, pzb, lb, yeb, ywb);
      syrk(yea, ywa);
      syrk(yeb, ywb);
      if (bad) status |= ISV_W_SINGULAR;
    }
    qx += 128; qy += 128; ql += 128;
    if (!ZONE) qz += 128;
  }
  // Tail: fewer than 128 landmarks left (already in the ring), predicated per lane.
  if (base < L) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i0 = base + 64 * h;
      if (i0 >= L) break;
      double yea[6], ywa[6], yeb[6], ywb[6];
      if (i0 + 32 + lane < L) {
        int bad = chain(rx[2 * h], ry[2 * h], rz[2 * h], rl[2 * h], yea, ywa);
        bad |= chain(rx[2 * h + 1], ry[2 * h + 1], rz[2 * h + 1], rl[2 * h + 1], yeb, ywb);
        syrk(yea, ywa);
        syrk(yeb, ywb);
        if (bad) status |= ISV_W_SINGULAR;
      } else if (i0 + lane < L) {
        if (chain(rx[2 * h], ry[2 * h], rz[2 * h], rl[2 * h], yea, ywa)) status |= ISV_W_SINGULAR;
        syrk(yea, ywa);
      }
    }
  }
  // cross-lane reduction through shared memory: lane l parks its 21 partial sums in column l of a
  // [21][33] tile, then lane t adds up row t (conflict-free both ways) -- 4x fewer instructions than
  // 42 shuffle trees.  Fixed summation order: results do not depend on scheduling.
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int t = 0; t < 21; ++t) R[t * kAccLd + lane] = half ? aw[t] : ae[t];
    __syncwarp();
    if (lane < 21) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += R[lane * kAccLd + j];
      g[half * 21 + lane] = acc;
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) status |= __shfl_xor_sync(kFullMask, status, o);
  if (lane == 0 && status && wstatus) atomicOr(wstatus, status);
}

// `next` != null: PERSISTENT form (ISV_TUNE_ACC_PERSIST warps per SM pull windows from a counter).  An experiment, off by
// default: 8 one-window CTAs per SM fill the register file, so nothing runs next to the landmark phase although it leaves
// 40 % of the FP64 pipe idle; 4 persistent warps per SM keep 70 % of the 8-warp throughput
// (profiles/r02r_occupancy_experiment.txt) and leave half of the registers to the backward kernel's CTAs.  Measured
// (profiles/r02u_persistent_accum_sweep.txt): the step does NOT get shorter -- 0.353 -> 0.388 ms at 4 warps per SM,
// 0.361 at 6; landmark and backward kernel take as long side by side as one after the other, with the same
// shared-memory carveout on every kernel and with 32 hardware queues as well.
template <bool ZONE, bool ISO, bool XYF = false>
__global__ void ISV_ACC_BOUNDS
marg_forward_accum_kernel(isv_batch_in in, double* __restrict__ gram, int32_t* wstatus, DevCfg cfg, int32_t* next) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  int win = blockIdx.x * kAccWarps + warp;
  for (;;) {
    if (next) {
      if (lane == 0) win = atomicAdd(next, 1);
      win = __shfl_sync(kFullMask, win, 0);
    }
    if (win >= in.n_windows) return;
    const long long lm0 = in.lm_offset[win];
    const int L = (int)(in.lm_offset[win + 1] - lm0);
    forward_accum_body<ZONE, ISO, XYF>(in, cfg, win, lane, smem + warp * kAccSmemPerWarp, lm0, L, gram + (size_t)win * 42,
                                       wstatus ? wstatus + win : nullptr);
    if (!next) return;
    __syncwarp();
  }
}

// ---- kernel 2: everything after the landmark sums (12x12 / 6x6 algebra), one warp per window ----
// Body for one warp and window.  K = kFwdSmemPerWarp doubles of shared memory owned by the warp; gw = the window's 42 Gram
// entries, F = its factor-Jacobian record (global in the batch kernel, shared in the fused one).
__device__ __forceinline__ void forward_tail_body(const isv_batch_in& in, const isv_batch_out& out, const DevCfg& cfg, const int win,
                                                  const int lane, double* K, const double* __restrict__ gw,
                                                  const double* __restrict__ F, int32_t* wstatus,
                                                  double* __restrict__ dbg_lamda_prior) {
  double* X = K + kFwdConst;                  // tail work;  K = Psi (12x6 column-major)
  int status = 0;
  int nonfinite = 0;

  const double* pose0 = in.pose_fwd + (size_t)win * 14;
  const double* pose1 = pose0 + 7;
  // every global load of this kernel is issued here, up front: the factor-Jacobian record (252 doubles, 8
  // per lane, parked in registers until the work area is free) and the Gram triangles
  double fjr[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) fjr[t] = (32 * t + lane < 252) ? F[32 * t + lane] : 0.0;
  // work map (doubles): S12[0] H12[144] Ye[288] Ys[324] tmp[360..432) ; then
  // Wst[288] G[432] Jr6[504] tA[612] tB[684] wk[756]
  double* S12 = X;
  double* H12 = X + 144;
  double* Ye = X + 288;   // 6 x 6 Gram matrices (ld 6)
  double* Ys = X + 324;
  double* Tm = X + 360;   // 12 x 6
  if (lane == 0) {
    Quat Qi = quat_from_pose(pose0), Qj = quat_from_pose(pose1);
    double R0[9], R1[9], D[9], dv[3], t3[3];
    q2R(Qi, R0);
    q2R(Qj, R1);
    mat3_tmul(R0, R1, D);
    for (int i = 0; i < 3; ++i) t3[i] = pose0[i] - pose1[i];
    mat3_tvec(R1, t3, dv);
    double* Psi = K;
    for (int i = 0; i < 72; ++i) Psi[i] = 0.0;
    double Sd[9];
    skew3(dv, Sd);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        Psi[r + 12 * (3 + c)] = -R1[3 * r + c];
        Psi[(3 + r) + 12 * c] = (r == c) ? -1.0 : 0.0;
        Psi[(3 + r) + 12 * (3 + c)] = -Sd[3 * r + c];
        Psi[(6 + r) + 12 * (3 + c)] = R1[3 * r + c];
        Psi[(9 + r) + 12 * c] = D[3 * r + c];
      }
  } else if (lane >= 8 && lane < 29) {
    // unpack the lower triangles: entry t = (i, j), j <= i, row-wise
    const int t = lane - 8;
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    const int j = t - i * (i + 1) / 2;
    const double ve = gw[t], vw = gw[21 + t];
    Ye[i + 6 * j] = ve; Ye[j + 6 * i] = ve;
    Ys[i + 6 * j] = vw; Ys[j + 6 * i] = vw;
  }
  __syncwarp();
  {
    const double* Psi = K;
    w_gemm_t<false, false, 12, 6, 6>(Psi, 12, Ys, 6, Tm, 12, 0, lane);
    w_gemm_t<false, true, 12, 12, 6>(Tm, 12, Psi, 12, S12, 12, 0, lane);   // Schur over the landmarks
    w_gemm_t<false, false, 12, 6, 6>(Psi, 12, Ye, 6, Tm, 12, 0, lane);
    w_gemm_t<false, true, 12, 12, 6>(Tm, 12, Psi, 12, H12, 12, 0, lane);   // sum e e^T
  }
  double* Wst = X + 288;
  double* G = X + 432;
  double* Jr6 = X + 504;
  double* tA = X + 612;
  double* tB = X + 684;
  double* wk = X + 756;
  double* o_se3 = out.se3_out + (size_t)win * ISV_SE3_REC;
  double* o_pg = out.pg_out + (size_t)win * ISV_PG_REC;
  // Wst (12x12: rows 0-5 = sp [0 | Jp], rows 6-11 = sr [Jj | Ji]), G (6x12) and Jr6 (6x6) come from
  // marg_factor_jac_kernel; they are contiguous in the scratch in the order of the work map
#pragma unroll
  for (int t = 0; t < 8; ++t)
    if (32 * t + lane < 252) Wst[32 * t + lane] = fjr[t];
  __syncwarp();
  // S12 += Wst^T Wst ; H12 = E12 + S12  (= Lamda[0:12,0:12], :1243)
  w_gemm_t<true, false, 12, 12, 12>(Wst, 12, Wst, 12, S12, 12, 1, lane);
  for (int idx = lane; idx < 144; idx += 32) H12[idx] += S12[idx];
  __syncwarp();
  // ---- pose-graph relative-pose factor (:1243-1259) -------------------------------------------
  // J = G (6x12, [Ji|Jj]) ; Jpinv = J^T (J J^T)^-1 (full row rank) ; rpOmega = Jpinv^T H12 Jpinv
  w_gemm_t<false, true, 6, 6, 12>(G, 6, G, 6, tA, 6, 0, lane);          // tA = J J^T
  if (w_spd_inverse_regs<6>(tA, 6, lane)) status |= ISV_W_SINGULAR;
  w_gemm_t<true, false, 12, 6, 6>(G, 6, tA, 6, tB, 12, 0, lane);        // tB = Jpinv (12x6)
  w_gemm_t<false, false, 12, 6, 12>(H12, 12, tB, 12, wk, 12, 0, lane);  // wk = H12 Jpinv
  w_gemm_t<true, false, 6, 6, 12>(tB, 12, wk, 12, tA, 6, 0, lane);      // tA = rpOmega
  w_copy(Wst, tA, 36, lane);
  if (w_llt_upper_regs<6>(Wst, 6, o_pg + 12, lane, nonfinite)) status |= ISV_W_NOT_SPD;
  w_copy(Wst, tA, 36, lane);
  if (w_spd_inverse_regs<6>(tA, 6, lane)) {                            // covRel = rpOmega^-1
    w_copy(tA, Wst, 36, lane);
    if (w_inverse(tA, 6, 6, wk, lane)) status |= ISV_W_SINGULAR;
  }
  for (int i = lane; i < 36; i += 32) {
    if (!isfinite(tA[i])) nonfinite = 1;
    o_pg[48 + i] = tA[i];
  }
  __syncwarp();
  // ---- Schur complement over T0 (:1286-1288 with the landmarks already eliminated) -------------
  w_copy2d(tA, 6, S12 + 6 + 12 * 6, 12, 6, 6, lane);                  // tA = S[6:12,6:12]
  if (w_spd_inverse_regs<6>(tA, 6, lane)) {
    w_copy2d(tA, 6, S12 + 6 + 12 * 6, 12, 6, 6, lane);
    if (w_inverse(tA, 6, 6, wk, lane)) status |= ISV_W_SINGULAR;
  }
  w_gemm_t<false, false, 6, 6, 6>(S12 + 12 * 6, 12, tA, 6, tB, 6, 0, lane);        // tB = S[0:6,6:12] Smm^-1
  w_copy2d(Wst, 6, S12, 12, 6, 6, lane);                                          // Wst = S[0:6,0:6]
  w_gemm_t<false, true, 6, 6, 6>(tB, 6, S12 + 12 * 6, 12, Wst, 6, -1, lane);       // Lamda_prior (6x6)
  // ---- rank decision + recovery of the SE3 prior on T1 (:1304-1349) ---------------------------
  // Fast path: Lamda_prior^-1 by the SPD route; when ||A||_F ||A^-1||_F < 1e10 every pivot of
  // Eigen's FullPivHouseholderQR is > 5e-13 of the largest, far above both its early-exit test
  // (6 eps) and the 1e-16 rank threshold, i.e. qr.rank() == 6 without running the QR.  Otherwise
  // the QR is restated literally on one lane.
  w_copy(tA, Wst, 36, lane);
  if (dbg_lamda_prior)   // forensic store: Lamda_prior (:1288) of the structured route, column-major 6 x 6
    for (int i = lane; i < 36; i += 32) dbg_lamda_prior[(size_t)win * 36 + i] = Wst[i];
  int rank = 0;
  {
    double fa = 0.0, fi = 0.0;
    for (int i = lane; i < 36; i += 32) fa = fma(tA[i], tA[i], fa);
    const int bad = w_spd_inverse_regs<6>(tA, 6, lane);
    for (int i = lane; i < 36; i += 32) fi = fma(tA[i], tA[i], fi);
    fa = warp_sum(fa);
    fi = warp_sum(fi);
    if (!bad && fa * fi < 1e20) rank = 6;
  }
  if (rank != 6) {
    if (lane == 0) {
      for (int i = 0; i < 36; ++i) tB[i] = Wst[i];
      rank = serial_fullpiv_qr_inverse<6>(tB, tA, cfg.qr_threshold);   // tA = cov = qr.solve(I)
    }
    rank = __shfl_sync(kFullMask, rank, 0);
    __syncwarp();
  }
  int out_rank = rank;
  if (rank == 6) {
    w_gemm_t<false, false, 6, 6, 6>(Jr6, 6, tA, 6, tB, 6, 0, lane);     // Jr cov
    w_gemm_t<false, true, 6, 6, 6>(tB, 6, Jr6, 6, tA, 6, 0, lane);      // covi = Jr cov Jr^T
  } else {
    status |= ISV_W_RANK_DEFICIENT;
    // truncated eigen path (:1311-1331), factored form: Lamda_prior = sum_k g_k g_k^T (rows of tB)
    for (int i = lane; i < 36; i += 32) tB[i] = 0.0;
    __syncwarp();
    const int nrow = w_pivoted_cholesky_rows(Wst, 6, 6, tB, 6, wk, lane);
    if (w_onesided_jacobi_rows<4, 2>(tB, 6, nrow, 6, wk, lane) >= 30) status |= ISV_W_EIG_NOCONV;
    int er = 0;
    for (int k = 0; k < nrow; ++k) er += (wk[k] > cfg.alpha) ? 1 : 0;
    out_rank = er;
    for (int idx = lane; idx < 36; idx += 32) {   // wk+8 = Jr G^T (6 x nrow)
      int r = idx % 6, k = idx / 6;
      double acc = 0.0;
      for (int c = 0; c < 6; ++c) acc = fma(Jr6[r + 6 * c], tB[c + 6 * k], acc);
      wk[8 + idx] = acc;
    }
    __syncwarp();
    for (int idx = lane; idx < 36; idx += 32) {
      int r = idx % 6, c = idx / 6;
      double acc = 0.0;
      for (int k = 0; k < nrow; ++k) {
        double lamk = wk[k];
        if (lamk > cfg.alpha) acc += wk[8 + r + 6 * k] * wk[8 + c + 6 * k] / (lamk * lamk);
      }
      tA[idx] = acc;
    }
    __syncwarp();
  }
  // sqrt_info = LLT(covi.inverse()).matrixL().transpose()  (:1349)
  if (w_sqrt_info_from_cov_regs<6>(tA, 6, o_se3 + 12, lane, nonfinite)) status |= ISV_W_NOT_SPD | ISV_W_SINGULAR;
  if (__any_sync(kFullMask, nonfinite)) status |= ISV_W_NONFINITE;
  // merge per-lane status bits
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) status |= __shfl_xor_sync(kFullMask, status, o);
  if (lane == 0) {
    out.rank[2 * win] = out_rank;
    if (wstatus) atomicOr(wstatus, status);
  }
}

__global__ void __launch_bounds__(kThreads, ISV_FWD_TAIL_MINB)
marg_forward_tail_kernel(isv_batch_in in, isv_batch_out out, const double* __restrict__ gram,
                         const double* __restrict__ fj, DevCfg cfg, double* __restrict__ dbg_lamda_prior) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int win = blockIdx.x * kWarpsPerCta + warp;
  if (win >= in.n_windows) return;
  forward_tail_body(in, out, cfg, win, lane, smem + warp * kFwdSmemPerWarp, gram + (size_t)win * 42, fj + (size_t)win * kFJ,
                    out.status ? out.status + win : nullptr, dbg_lamda_prior);
}

// =================================================================================================
// MargBackward -- square-root-information form.
//   M (24 x 30) = [ s_vb on VB_{V-1} ; L^-1 J_imu ]   with  covariance = L L^T   (rows 0-8 / 9-23)
//   Householder-eliminate the VB_{V-1} columns -> rows 9-23, columns 0-20 = G (15 x 21) with
//   G^T G = Lamda_prior (:1419);  one-sided Jacobi makes the rows of G orthogonal:
//   Lamda_prior = sum_k g_k g_k^T, eigenvalues |g_k|^2 (:1479-1497), and for every recovered factor
//   cov_i = J_i U D^-1 U^T J_i^T = sum_{|g_k|^2 > ALPHA} (J_i g_k)(J_i g_k)^T / |g_k|^4  (:1500-1516).
// Algebraically identical to the reference's information-form route, without its cancellation.
// =================================================================================================
// Column-per-lane formulation: lane c < 30 owns column c of M in REGISTERS from its creation to the
// end of the elimination (rows 9-23: 15 doubles; lanes 21-29 also the 9 prior rows).  Only the
// Cholesky columns of the covariance and the current Householder vector travel through shared memory
// (broadcast loads), so the elimination costs one LDS + two DFMA per element instead of four LDS, one
// STS and two DFMA.  The prior's sqrt_info is upper triangular (LLT(...).matrixL().transpose()), so
// reflector k only touches row k and the 15 IMU rows.
constexpr int kGld = 21;  // G rows in shared memory: element (k, c) at Gs[k * 21 + c]
constexpr int kBwdLc = 0, kBwdHv = 272, kBwdPv = 296, kBwdGs = 344, kBwdT = 660;
constexpr int kLcLd = 16;   // Cholesky factor of the covariance: column k at Lc[16 k + i] (16-byte aligned pairs), [256..271) = 1 / L_kk
// Named barrier of the fused single-event kernel: the backward warp waits here (after the covariance Cholesky, which needs
// only the pre-integration record) for the warps that evaluate the IMU / relative-pose / roll-pitch Jacobians.
constexpr int kFusedBarBwd = 1, kFusedBarFwd = 2;
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Body for one warp and window.  S = kBwdSmemPerWarp doubles of shared memory owned by the warp, F = the window's
// factor-Jacobian record.  FUSED_BAR > 0: F is produced concurrently by other warps of the CTA; it is not touched before the
// named barrier kFusedBarBwd (FUSED_BAR threads) that follows the covariance Cholesky.
template <int FUSED_BAR>
__device__ __forceinline__ void backward_body(const isv_batch_in& in, const isv_batch_out& out, const DevCfg& cfg, const int win,
                                              const int lane, double* S, const double* __restrict__ F, int32_t* wstatus,
                                              double* __restrict__ dbg_g, long long* clk = nullptr) {
  // profiling aid of the fused kernel (isv_test_fused_stamps): cycle counter at the phase boundaries
  int n_clk = 0;
  auto tick = [&]() {
    if (FUSED_BAR > 0 && clk && lane == 0) clk[n_clk] = clock64();
    ++n_clk;
  };
  double* Lc = S + kBwdLc;     // Cholesky factor of the covariance, column k at Lc[16 k + i]; [256..271) = 1 / L_kk
  double* hv = S + kBwdHv;     // current Householder vector (15) + tau
  double* pvt = S + kBwdPv;    // LQ: the pivot row of the current step (two buffers of 24, alternating)
  double* Gs = S + kBwdGs;     // 15 x 21
  double* T = S + kBwdT;       // Jrel (72) | Jrp (16) | JU / Ls (256) | cov 9x9, 6x6, 2x2 (128)
  double* sc = T + 472;        // scratch of the general path: dinv[0..21), lam[24..)
  int status = 0, nonfinite = 0;

  const double* pvb = in.prior_vb + (size_t)win * ISV_VB_REC;
  const double* pre = in.preint + (size_t)win * ISV_PREINT_REC;
  double* o_rel = out.rel_out + (size_t)win * ISV_REL_REC;
  double* o_vb = out.vb_out + (size_t)win * ISV_VB_REC;
  double* o_rp = out.rp_out + (size_t)win * ISV_RP_REC;
  double* Jrel = T;          // Ji (36) | Jj (36)
  double* Jrp = T + 72;      // 2 x 6
  double* JU = T + 88;       // 17 x 15 (fast path: Y) / up to 9 x 15 (eigen path)
  double* cov = T + 344;     // up to 9 x 9
  // [Ji | Jj | Jrp] of the recovered factors: loaded now, parked in shared memory after the Cholesky (the
  // global-load latency hides behind it)
  double jr0 = 0.0, jr1 = 0.0, jr2 = 0.0;
  if (FUSED_BAR == 0) { jr0 = F[kFJ_REL + lane]; jr1 = F[kFJ_REL + 32 + lane]; jr2 = (lane < 20) ? F[kFJ_REL + 64 + lane] : 0.0; }

  // ---- covariance = L L^T, one column per lane in registers (right-looking) ------------------------
  {
    const int j = lane < 15 ? lane : 14;
    double a[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) a[i] = pre[17 + 225 + i + 15 * j];
#pragma unroll
    for (int k = 0; k < 15; ++k) {
      if (lane == k) {
        const double d = a[k];
        if (!(d > 0.0)) status |= ISV_W_NOT_SPD;
        const double ri = fast_rsqrt(d);
        Lc[256 + k] = ri;
        // column k, rows k..14, stored as aligned pairs (row k - 1 of an odd k rides along: never read)
        double2* c2 = reinterpret_cast<double2*>(Lc + kLcLd * k);
#pragma unroll
        for (int i = k & ~1; i < 16; i += 2) {
          const double lo = (i == k) ? d * ri : a[i < 15 ? i : 14] * ri;
          const double hi = (i + 1 == k) ? d * ri : a[i + 1 < 15 ? i + 1 : 14] * ri;
          c2[i >> 1] = make_double2(lo, hi);
        }
      }
      __syncwarp();
      if (lane > k && lane < 15) {
        const double ljk = Lc[kLcLd * k + j];
        const double2* c2 = reinterpret_cast<const double2*>(Lc + kLcLd * k);
        // rows < j are dead weight; row k itself (when the first pair starts at k) is dead after its pivot
#pragma unroll
        for (int i = (k + 1) & ~1; i < 15; i += 2) {
          const double2 l2 = c2[i >> 1];
          if (i > k) a[i] = fma(-l2.x, ljk, a[i]);
          if (i + 1 < 15) a[i + 1] = fma(-l2.y, ljk, a[i + 1]);
        }
      }
    }
    __syncwarp();
  }
  // ---- my column: IMU rows (unweighted Jacobian from marg_factor_jac_kernel, row-major 15 x 30:
  //      coalesced) and, for the VB_{V-1} columns, the prior rows (sqrt_info column, :1372-1380) -----
  tick();   // [0] covariance Cholesky done
  if (FUSED_BAR > 0) {
    named_bar_sync(kFusedBarBwd, FUSED_BAR);
    tick();   // [1] factor Jacobians arrived
    jr0 = F[kFJ_REL + lane]; jr1 = F[kFJ_REL + 32 + lane]; jr2 = (lane < 20) ? F[kFJ_REL + 64 + lane] : 0.0;
  }
  const int c = lane < 30 ? lane : 29;
  double col[15], pr[9];
#pragma unroll
  for (int i = 0; i < 15; ++i) col[i] = F[kFJ_IMU + 30 * i + c];
#pragma unroll
  for (int i = 0; i < 9; ++i) pr[i] = (lane >= 21 && lane < 30) ? pvb[9 + i + 9 * (lane - 21)] : 0.0;

  T[lane] = jr0;
  T[32 + lane] = jr1;
  if (lane < 20) T[64 + lane] = jr2;
  // ---- sqrt_info^T sqrt_info = covariance^-1 (imu_factor.h:181): col <- L^-1 col ------------------
#pragma unroll
  for (int i = 0; i < 15; ++i) {
    double sacc = col[i];
#pragma unroll
    for (int l = 0; l < i; ++l) sacc = fma(-Lc[kLcLd * l + i], col[l], sacc);
    col[i] = sacc * Lc[256 + i];
  }
  // ---- Schur complement over VB_{V-1} (:1413-1419) as a QR elimination of columns 21..29 -----------
  // Reflectors in unnormalised form  H = I - gamma w w^T ,  w = [x0 - beta ; tail] ,
  // gamma = 1 / (|beta| (|beta| + |x0|)) ,  |beta| = sqrt(x0^2 + |tail|^2) = nrm * rsqrt(nrm):
  // one rsqrt and one reciprocal per reflector instead of a sqrt and two divisions, and the tail is used
  // as it is (no scaling pass).
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (lane == 21 + k) {
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
      for (int i = 0; i < 15; i += 3) { t0 = fma(col[i], col[i], t0); t1 = fma(col[i + 1], col[i + 1], t1); t2 = fma(col[i + 2], col[i + 2], t2); }
      const double tt = (t0 + t1) + t2;
      const double x0 = pr[k];
      double gamma = 0.0, w0 = 0.0;
      if (tt > 0.0) {
        const double nrm = fma(x0, x0, tt);
        const double ab = nrm * fast_rsqrt(nrm);       // |beta|
        const double ax = fabs(x0);
        gamma = fast_rcp(ab * (ab + ax));
        w0 = (x0 >= 0.0) ? (ax + ab) : -(ax + ab);     // x0 - beta , beta = -sign(x0) |beta|
      }
      // the reflector travels as nine 16-byte words: [0..14] tail, [15] gamma, [16] w0 ([17] unused)
      double2* h2 = reinterpret_cast<double2*>(hv);
#pragma unroll
      for (int i = 0; i < 14; i += 2) h2[i >> 1] = make_double2(col[i], col[i + 1]);
      h2[7] = make_double2(col[14], gamma);
      h2[8] = make_double2(w0, 0.0);
    }
    __syncwarp();
    {
      // apply H to every live column other than 21+k; row k is zero outside the VB columns
      const double2* h2 = reinterpret_cast<const double2*>(hv);
      double hr[16];
#pragma unroll
      for (int i = 0; i < 16; i += 2) { const double2 t = h2[i >> 1]; hr[i] = t.x; hr[i + 1] = t.y; }
      const double gamma = hr[15], w0 = h2[8].x;
      double d0 = (lane > 21 + k && lane < 30) ? pr[k] * w0 : 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
      for (int i = 0; i < 15; i += 3) { d0 = fma(hr[i], col[i], d0); d1 = fma(hr[i + 1], col[i + 1], d1); d2 = fma(hr[i + 2], col[i + 2], d2); }
      const double sd = gamma * ((d0 + d1) + d2);
      if (lane < 21 || lane > 21 + k) {
#pragma unroll
        for (int i = 0; i < 15; ++i) col[i] = fma(-sd, hr[i], col[i]);
      }
    }
    __syncwarp();
  }
  tick();   // [2] whitening + 9 Householder steps done
  // rows 9-23 of columns 0..20 are G (15 x 21) with G^T G = Lamda_prior: transpose to one row per lane
  if (lane < 21) {
#pragma unroll
    for (int i = 0; i < 15; ++i) Gs[i * kGld + lane] = col[i];
    if (dbg_g)   // forensic store: G (15 x 21) with G^T G = Lamda_prior (:1419)
#pragma unroll
      for (int i = 0; i < 15; ++i) dbg_g[(size_t)win * 315 + i * 21 + lane] = col[i];
  }
  __syncwarp();
  // ---- fast path: no eigen-decomposition when every non-zero eigenvalue is provably > ALPHA ----
  // LQ-factorise G = [L 0] Q (Householder from the right, one ROW per lane in registers) and carry
  // the 17 rows of the recovered-factor Jacobians Jr (:1456-1464) through the same reflectors:
  //   Lamda_prior = Q^T [L^T L 0; 0 0] Q ,  pinv = Q^T [(L^T L)^-1 0; 0 0] Q ,
  //   cov_i = J_i pinv J_i^T = Y_i Y_i^T  with  y_r = L^-T (Jr Q^T)_r[0:15] .
  // sigma_min(L)^2 >= 1 / ||L^-1||_F^2 is a lower bound of the smallest non-zero eigenvalue; if it
  // exceeds ALPHA the reference's cut (:1482) keeps exactly these 15 and U D^-1 U^T == pinv.
  double row[21];
  if (lane < 15) {
#pragma unroll
    for (int cc = 0; cc < 21; ++cc) row[cc] = Gs[lane * kGld + cc];
  } else {
#pragma unroll
    for (int cc = 0; cc < 21; ++cc) row[cc] = 0.0;
    const int r = lane - 15;
    if (r < 6) {          // relative pose: Jj -> cols 0:6 (T_V), Ji -> cols 15:21 (T_{V-1})
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) { row[cc] = Jrel[36 + r + 6 * cc]; row[15 + cc] = Jrel[r + 6 * cc]; }
    } else if (r < 15) {  // speed-bias prior: I9 at cols 6:15
#pragma unroll
      for (int cc = 0; cc < 9; ++cc) row[6 + cc] = (cc == r - 6) ? 1.0 : 0.0;
    } else {              // roll/pitch: 2x6 at cols 15:21
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) row[15 + cc] = Jrp[(r - 15) + 2 * cc];
    }
  }
  // (reflectors in the unnormalised form of the elimination above: H = I - gamma w w^T, w = [x0 - beta ; x_tail])
  // The pivot row of step j is broadcast through shared memory: its owner stores the tail (and |tail|^2, x0) as
  // 16-byte words, everybody reads them back with 128-bit broadcast loads -- a quarter of the MIO instructions of the
  // 64-bit shuffles this replaced (ncu: the shuffles were MIO-throttled half of the time).  Two buffers alternate so
  // that one __syncwarp per step suffices.  A zero tail (tt == 0: the row is already reduced) gives gamma = 0, i.e. the
  // identity, without a branch: the steps stay in one basic block and overlap.
#pragma unroll
  for (int j = 0; j < 15; ++j) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int c = j + 1; c < 21; ++c) {
      if ((c - j) % 3 == 1) t0 = fma(row[c], row[c], t0);
      else if ((c - j) % 3 == 2) t1 = fma(row[c], row[c], t1);
      else t2 = fma(row[c], row[c], t2);
    }
    double2* pw = reinterpret_cast<double2*>(pvt + 24 * (j & 1));
    constexpr int kFirstPair = 0;   // (placeholder so the pair index below reads naturally)
    if (lane == j) {
#pragma unroll
      for (int pp = (j + 1) >> 1; pp < 11; ++pp) pw[pp] = make_double2(row[2 * pp], (2 * pp + 1 < 21) ? row[2 * pp + 1] : 0.0);
      pw[11] = make_double2((t0 + t1) + t2, row[j]);
    }
    __syncwarp();
    const double2 hx = pw[11];
    const double tt = hx.x, x0 = hx.y;
    const bool nzt = tt > 0.0;
    const double nrm = fma(x0, x0, tt);
    const double ab = nrm * fast_rsqrt(nrm);         // |beta|
    const double ax = fabs(x0);
    const double gamma = nzt ? fast_rcp(ab * (ab + ax)) : 0.0;
    const double w0 = nzt ? ((x0 >= 0.0) ? (ax + ab) : -(ax + ab)) : 0.0;   // x0 - beta
    const double beta = nzt ? ((x0 >= 0.0) ? -ab : ab) : x0;
    double v[22];
#pragma unroll
    for (int pp = (j + 1) >> 1; pp < 11; ++pp) { const double2 t = pw[pp]; v[2 * pp] = t.x; v[2 * pp + 1] = t.y; }
    double d0 = row[j] * w0, d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int c = j + 1; c < 21; ++c) {
      if ((c - j) % 3 == 1) d0 = fma(v[c], row[c], d0);
      else if ((c - j) % 3 == 2) d1 = fma(v[c], row[c], d1);
      else d2 = fma(v[c], row[c], d2);
    }
    const double sdot = gamma * ((d0 + d1) + d2);
    // uniform update (no divergence): the pivot lane's own tail becomes rounding-level garbage instead of
    // exact zeros -- it is never read again (later reflectors read lane j' > j, the solve reads only the
    // lower triangle of L)
    row[j] = (lane == j) ? beta : fma(-sdot, w0, row[j]);
#pragma unroll
    for (int c = j + 1; c < 21; ++c) row[c] = fma(-sdot, v[c], row[c]);
    (void)kFirstPair;
  }
  __syncwarp();
  tick();   // [3] LQ done
  // L (rows of lanes 0-14, lower triangular) -> shared; solve L^T y = rhs for all 32 lanes at once:
  // lanes 0-14: rhs = e_lane (columns of L^-T, for the eigenvalue bound), lanes 15-31: rhs = (Jr Q^T)_r
  double* Ls = T + 88;  // 15 x 15 (ld 16: column k at Ls[16 k + m], 16-byte aligned pairs) + 15 reciprocal diagonals at [240..);
                        // overlays JU, which is written later
  if (lane < 15) {
#pragma unroll
    for (int c = 0; c < 15; ++c) Ls[lane + 16 * c] = row[c];
    Ls[240 + lane] = fast_rcp(row[lane]);
  }
  __syncwarp();
  double y[16];
  y[15] = 0.0;
#pragma unroll
  for (int k = 14; k >= 0; --k) {
    double sacc = (lane < 15) ? ((lane == k) ? 1.0 : 0.0) : row[k];
    const double2* l2 = reinterpret_cast<const double2*>(Ls + 16 * k);
#pragma unroll
    for (int m2 = (k + 1) & ~1; m2 < 15; m2 += 2) {
      const double2 t = l2[m2 >> 1];
      if (m2 > k) sacc = fma(-t.x, y[m2], sacc);
      if (m2 + 1 < 15) sacc = fma(-t.y, y[m2 + 1], sacc);
    }
    y[k] = sacc * Ls[240 + k];
  }
  double ninv2 = 0.0;
#pragma unroll
  for (int k = 0; k < 15; ++k) ninv2 = fma(y[k], y[k], ninv2);
  ninv2 = (lane < 15) ? ninv2 : 0.0;
  ninv2 = warp_sum(ninv2);
  int rank;
  const bool fast = isfinite(ninv2) && ninv2 > 0.0 && (1.0 / ninv2) > cfg.alpha;
  tick();   // [4] back-substitution + eigenvalue bound done
  if (fast) {
    rank = 15;
    __syncwarp();
    if (lane >= 15) {
#pragma unroll
      for (int k = 0; k < 15; ++k) JU[(lane - 15) + 17 * k] = y[k];
    }
    __syncwarp();
    // cov blocks: rel = rows 0-5, vb = rows 6-14, rp = rows 15-16 of Y ; cov_i = Y_i Y_i^T
    double* cov6 = cov + 81;
    double* cov2 = cov + 117;
    for (int idx = lane; idx < 121; idx += 32) {
      int r0, n, e;
      double* dst;
      if (idx < 81) { r0 = 6; n = 9; e = idx; dst = cov; }
      else if (idx < 117) { r0 = 0; n = 6; e = idx - 81; dst = cov6; }
      else { r0 = 15; n = 2; e = idx - 117; dst = cov2; }
      const int r = e % n, c = e / n;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 15; ++k) acc = fma(JU[r0 + r + 17 * k], JU[r0 + c + 17 * k], acc);
      dst[e] = acc;
    }
    __syncwarp();
    tick();   // [5] covariance blocks done
    {
      // the three sqrt-information factors at once: lanes 0-8 vb (9x9), 9-14 rel (6x6), 15-16 rp (2x2);
      // the per-group scratch overlays JU, which is dead from here on
      const bool act = lane < 17;
      const int grp = lane < 9 ? 0 : (lane < 15 ? 1 : 2);
      const int N = grp == 0 ? 9 : (grp == 1 ? 6 : 2);
      const int c = grp == 0 ? lane : (grp == 1 ? lane - 9 : lane - 15);
      const double* Ag = grp == 0 ? cov : (grp == 1 ? cov6 : cov2);
      double* Usg = JU + (grp == 0 ? 0 : (grp == 1 ? 90 : 132));
      double* og = grp == 0 ? o_vb + 9 : (grp == 1 ? o_rel + 12 : o_rp + 9);
      if (w_sqrt_info_multi<9>(Ag, N, c, act, Usg, og, nonfinite)) status |= ISV_W_NOT_SPD;
    }
  } else {
    // ---- general path: eigen-decomposition (:1479-1497) by one-sided Jacobi on the rows of G ----
    if (w_onesided_jacobi_rows<4, 6>(Gs, kGld, 15, 21, sc + 24, lane, 30, 1) >= 30) status |= ISV_W_EIG_NOCONV;
    int keepf = 0;
    if (lane < 21) {
      double lamk = (lane < 15) ? sc[24 + lane] : 0.0;
      int keep = lamk > cfg.alpha;  // strict, Q12
      sc[lane] = keep ? 1.0 / (lamk * lamk) : 0.0;
      keepf = keep;
    }
    rank = __popc(__ballot_sync(kFullMask, keepf));
    __syncwarp();
    const double* dinv = sc;
    for (int idx = lane; idx < 6 * 15; idx += 32) {
      int r = idx % 6, k = idx / 6;
      double acc = 0.0;
      for (int c = 0; c < 6; ++c) {
        acc = fma(Jrel[36 + r + 6 * c], Gs[k * kGld + c], acc);
        acc = fma(Jrel[r + 6 * c], Gs[k * kGld + 15 + c], acc);
      }
      JU[idx] = acc;
    }
    __syncwarp();
    for (int idx = lane; idx < 36; idx += 32) {
      int r = idx % 6, c = idx / 6;
      double acc = 0.0;
      for (int k = 0; k < 15; ++k) acc = fma(JU[r + 6 * k] * dinv[k], JU[c + 6 * k], acc);
      cov[idx] = acc;
    }
    __syncwarp();
    if (w_sqrt_info_from_cov_regs<6>(cov, 6, o_rel + 12, lane, nonfinite)) status |= ISV_W_NOT_SPD;
    for (int idx = lane; idx < 81; idx += 32) {
      int r = idx % 9, c = idx / 9;
      double acc = 0.0;
      for (int k = 0; k < 15; ++k) acc = fma(Gs[k * kGld + 6 + r] * dinv[k], Gs[k * kGld + 6 + c], acc);
      cov[idx] = acc;
    }
    __syncwarp();
    if (w_sqrt_info_from_cov<9>(cov, 9, o_vb + 9, lane, nonfinite)) status |= ISV_W_NOT_SPD;
    for (int idx = lane; idx < 2 * 15; idx += 32) {
      int r = idx % 2, k = idx / 2;
      double acc = 0.0;
      for (int c = 0; c < 6; ++c) acc = fma(Jrp[r + 2 * c], Gs[k * kGld + 15 + c], acc);
      JU[idx] = acc;
    }
    __syncwarp();
    for (int idx = lane; idx < 4; idx += 32) {
      int r = idx % 2, c = idx / 2;
      double acc = 0.0;
      for (int k = 0; k < 15; ++k) acc = fma(JU[r + 2 * k] * dinv[k], JU[c + 2 * k], acc);
      cov[idx] = acc;
    }
    __syncwarp();
    if (w_sqrt_info_from_cov_regs<2>(cov, 2, o_rp + 9, lane, nonfinite)) status |= ISV_W_NOT_SPD;
  }
  if (__any_sync(kFullMask, nonfinite)) status |= ISV_W_NONFINITE;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) status |= __shfl_xor_sync(kFullMask, status, o);
  if (lane == 0) {
    out.rank[2 * win + 1] = rank;
    if (wstatus) atomicOr(wstatus, status);
  }
}

__global__ void __launch_bounds__(kThreads, ISV_BWD_MINB)
marg_backward_kernel(isv_batch_in in, isv_batch_out out, const double* __restrict__ fj, DevCfg cfg, int vo_size,
                     double* __restrict__ dbg_g) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int win = blockIdx.x * kWarpsPerCta + warp;
  if (win >= in.n_windows) return;
  backward_body<0>(in, out, cfg, win, lane, smem + warp * kBwdSmemPerWarp, fj + (size_t)win * kFJ,
                   out.status ? out.status + win : nullptr, dbg_g);
  (void)vo_size;
}

}  // namespace isv
