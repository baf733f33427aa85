// Generic marginalization engine behind the `MarginalizationInfo` facade the north_star names
// (addResidualBlockInfo / preMarginalize / marginalize / getParameterBlocks).  IS-VINS itself deleted
// VINS-Mono's marginalization_factor.{h,cpp}, so this follows the PUBLISHED
// VINS-Mono algorithm (HKUST-Aerial-Robotics/VINS-Mono, vins_estimator/src/factor/
// marginalization_factor.cpp, `MarginalizationInfo::marginalize`): parity is unpinned by the reference.
//
//   ne_build_kernel        ThreadsConstructA: A += J_i^T J_j , b += J_i^T r over all residual blocks,
//                          scattered by parameter-block position -- one warp per residual block, its
//                          Jacobians staged in shared memory, every upper-triangle product formed once and
//                          reduced into A (and its mirror) with FP64 atomics
//   marg_schur_eig_kernel  A_rr - A_rm A_mm^+ A_mr ,  b_rr - A_rm A_mm^+ b_mm  (eigen-thresholded pseudo-inverse of
//                          the dense marginalized block) -- one CTA per problem
//   sym_eig_prior_kernel   (isv_sym_eig.cuh) eigen-decomposition of the reduced system,
//                          linearized_jacobians = S^1/2 V^T , residuals = S^-1/2 V^T b
// Layout of the tangent vector: [ m_dense | m_diag | n_keep ].  `m_diag` marginalized blocks are
// scalars that no residual block couples with each other (inverse depths: every ProjectionFactor
// touches exactly one), so that part of A_mm is diagonal and its elimination is the dense product
//   C -= X D^-1 X^T ,  X = A[R, diag]   (R = dense-marginalized + kept rows)
// which runs on the FP64 tensor cores (mma.sync.m8n8k4.f64 -> DMMA.8x8x4) with full 8x8 tiles: the one
// place on this path where the tensor pipe beats DFMA (isv_window_kernels.cuh explains why the 6x6
// landmark Grams do not).  The remaining m_dense x m_dense block (pose + speed-bias of the oldest
// frame, <= 32) gets VINS-Mono's eigen-thresholded pseudo-inverse.  Block-wise pseudo-inversion
// equals the joint one of VINS-Mono whenever A_mm has no eigenvalue <= eps; a scalar pivot or dense
// eigenvalue <= eps raises ISV_W_RANK_DEFICIENT.
#pragma once
#include "isv_device_math.cuh"
#include "isv_warp_linalg.cuh"
#include "isv_window_kernels.cuh"   // dmma884

#include "../../include/isv_capi.h"

namespace isv {

constexpr int kNeMaxRes = 15, kNeMaxCols = 36, kNeWarps = 4, kNeLd = kNeMaxCols + 1;
constexpr int kNeSmemPerWarp = kNeMaxRes * kNeLd + 16 + kNeMaxCols / 2 + 2;   // J | r | colpos (ints) in doubles

// One warp per residual block.  Its row-major Jacobian blocks are staged into shared memory as one
// n_res x tot matrix (ld 37: conflict-free down a column), together with the tangent position of every staged
// column; then every lane takes column PAIRS (a <= c) of the upper triangle -- decoded from a flat pair index, so all
// 32 lanes are busy whatever tot is -- forms J[:,a] . J[:,c] once and adds it to A at (pos(a), pos(c)) and at the
// mirrored entry with FP64 reductions (RED.ADD.F64).  The block -> position map is the caller's table: bit-exact.
__global__ void __launch_bounds__(32 * kNeWarps)
ne_build_kernel(isv_marg_generic_in in, double* __restrict__ A, double* __restrict__ b, int32_t* status) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long f = (long long)blockIdx.x * kNeWarps + warp;
  if (f >= in.n_factors) return;
  double* J = smem + warp * kNeSmemPerWarp;   // n_res x tot, row-major, ld = kNeLd
  double* r = J + kNeMaxRes * kNeLd;
  int* colpos = reinterpret_cast<int*>(r + 16);   // tangent position of every staged column
  const isv_ne_factor fa = in.factors[f];
  const int nres = fa.n_res, nb = fa.n_blocks;
  bool ok = nres >= 1 && nres <= kNeMaxRes && nb >= 1 && nb <= 5 && fa.problem >= 0 && fa.problem < in.n_problems;
  // lane k < nb owns block k: its record, and by a shuffle scan its first staged column
  isv_ne_block bl = {0, 0, 0, 0, 0};
  if (ok && lane < nb) {
    bl = in.blocks[fa.first_block + lane];
    if (bl.local_size < 1 || bl.local_size > 9 || bl.pos < 0 || bl.pos + bl.local_size > in.pos) ok = false;
  }
  ok = __all_sync(kFullMask, ok);
  int c0 = lane < nb ? bl.local_size : 0;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const int up = __shfl_up_sync(kFullMask, c0, o);
    if (lane >= o) c0 += up;
  }
  const int tot = __shfl_sync(kFullMask, c0, 7);    // inclusive scan: lanes >= nb repeat the total
  c0 -= lane < nb ? bl.local_size : 0;
  if (!ok || tot > kNeMaxCols) {
    if (lane == 0 && status) atomicOr(status, ISV_W_BAD_INDEX);
    return;
  }
  if (lane < nres) r[lane] = in.values[fa.res_offset + lane];
  for (int k = 0; k < nb; ++k) {
    const long long jo = __shfl_sync(kFullMask, bl.jac_offset, k);
    const int rs = __shfl_sync(kFullMask, bl.row_stride, k), ls = __shfl_sync(kFullMask, bl.local_size, k);
    const int ck = __shfl_sync(kFullMask, c0, k), pk = __shfl_sync(kFullMask, bl.pos, k);
    if (lane < ls) colpos[ck + lane] = pk + lane;
    // rows x local columns of the block; lanes along the columns first (ls <= 9: rows per pass = 32 / ls)
    const int rpp = 32 / ls, row_l = lane / ls, col_l = lane - row_l * ls;
    if (row_l < rpp)
      for (int row = row_l; row < nres; row += rpp)
        J[row * kNeLd + ck + col_l] = in.values[jo + (long long)row * rs + col_l];
  }
  __syncwarp();
  double* Ap = A + (size_t)fa.problem * in.pos * in.pos;
  double* bp = b + (size_t)fa.problem * in.pos;
  const size_t ldA = (size_t)in.pos;
  const int npair = tot * (tot + 1) / 2;
  const int dlo = in.m_dense, dhi = in.m_dense + in.m_diag;   // the diagonal marginalized block
  int coupled = 0;
  for (int p = lane; p < npair; p += 32) {
    // p = c (c + 1) / 2 + a with a <= c
    int c = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    if ((c + 1) * (c + 2) / 2 <= p) ++c;
    if (c * (c + 1) / 2 > p) --c;
    const int a = p - c * (c + 1) / 2;
    double acc = 0.0;
    for (int l = 0; l < nres; ++l) acc = fma(J[l * kNeLd + a], J[l * kNeLd + c], acc);
    if (acc != 0.0) {
      const int pa = colpos[a], pc = colpos[c];
      atomicAdd(Ap + pa + ldA * pc, acc);
      if (a != c) atomicAdd(Ap + pc + ldA * pa, acc);
      if (pa != pc && pa >= dlo && pa < dhi && pc >= dlo && pc < dhi) coupled = 1;
    }
  }
  // an off-diagonal entry inside the "diagonal" block: schur_diag_dmma_kernel reads only A[d, d] and would drop it
  if (__any_sync(kFullMask, coupled) && lane == 0 && status) atomicOr(status + fa.problem, ISV_W_DIAG_COUPLED);
  for (int a = lane; a < tot; a += 32) {
    double acc = 0.0;
    for (int l = 0; l < nres; ++l) acc = fma(J[l * kNeLd + a], r[l], acc);
    if (acc != 0.0) atomicAdd(bp + colpos[a], acc);
  }
}

// -------------------------------------------------------------------------------------------------
// Phase A as its own kernel: the Schur complement over the diagonal (inverse-depth) block,
//     C[R, R] -= X D^-1 X^T ,   b[R] -= X D^-1 b_diag ,   X = A[R, diag]  (nr x m_diag)
// as a tiled SYRK on the FP64 tensor cores.
// One CTA (4 warps) per 64 x 64 tile of the upper triangle
// of C and per problem; the K loop runs over the landmarks in slabs of 16 staged in shared memory
// (coalesced: for a fixed landmark the rows of X are contiguous in the column-major A); every warp owns a
// 32 x 32 patch = 4 x 4 DMMA.8x8x4 tiles (32 accumulator registers), 16 DMMA per 8 fragment loads; several
// CTAs per SM hide the slab loads of one behind the DMMAs of the others.  The
// slab leading dimension 68 makes the 64-bit fragment loads bank-conflict free (k * 68 mod 16 = 0,4,8,12).
// -------------------------------------------------------------------------------------------------
constexpr int kSdTile = 64, kSdSlab = 16, kSdLd = 68, kSdThreads = 128;

__global__ void __launch_bounds__(kSdThreads)
schur_diag_dmma_kernel(isv_marg_generic_in in, isv_marg_generic_out out) {
  __shared__ double Xa[kSdSlab * kSdLd];   // rows of tile ti, scaled by 1/d_l
  __shared__ double Xb[kSdSlab * kSdLd];   // rows of tile tj
  __shared__ double dinv_s[kSdSlab], bd_s[kSdSlab];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int prob = blockIdx.y;
  const int pos = in.pos, md = in.m_dense, mg = in.m_diag, m = md + mg, nr = pos - mg;
  const int nt = (nr + kSdTile - 1) / kSdTile;
  // blockIdx.x enumerates the upper-triangular tile pairs (ti <= tj)
  int ti = 0, rem = blockIdx.x;
  while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
  const int tj = ti + rem;
  double* A = out.A + (size_t)prob * pos * pos;
  double* b = out.b + (size_t)prob * pos;
  const double eps = in.eps;
  auto R = [&](int i) { return i < md ? i : i + mg; };
  const int fr = lane >> 2, fk = lane & 3;
  const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32;   // this warp's 32 x 32 patch inside the tile
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }
  double bacc = 0.0;
  // register double-buffering: the 16 global loads of slab s+1 (8 per operand and thread, all independent)
  // are issued before the DMMAs of slab s and parked in shared memory after them
  constexpr int kPer = kSdSlab * kSdTile / kSdThreads;   // 8
  double va[kPer], vb[kPer], dv = 0.0, bv = 0.0;
  auto fetch = [&](int l0) {
#pragma unroll
    for (int it = 0; it < kPer; ++it) {
      const int idx = tid + it * kSdThreads;
      const int k = idx / kSdTile, r = idx - k * kSdTile;
      const int l = l0 + k;
      const int ia = kSdTile * ti + r, ib = kSdTile * tj + r;
      va[it] = 0.0;
      vb[it] = 0.0;
      if (l < mg) {
        const size_t col = (size_t)pos * (md + l);
        if (ia < nr) va[it] = A[R(ia) + col];
        if (ib < nr) vb[it] = A[R(ib) + col];
      }
    }
    if (tid < kSdSlab) {
      const int l = l0 + tid;
      dv = 0.0;
      bv = 0.0;
      if (l < mg) { dv = A[(md + l) + (size_t)pos * (md + l)]; bv = b[md + l]; }
    }
  };
  fetch(0);
  for (int l0 = 0; l0 < mg; l0 += kSdSlab) {
    if (tid < kSdSlab) {
      dinv_s[tid] = dv > eps ? 1.0 / dv : 0.0;
      bd_s[tid] = bv;
      if (l0 + tid < mg && !(dv > eps) && blockIdx.x == 0 && out.status) atomicOr(out.status + prob, ISV_W_RANK_DEFICIENT);
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kPer; ++it) {
      const int idx = tid + it * kSdThreads;
      const int k = idx / kSdTile, r = idx - k * kSdTile;
      Xa[k * kSdLd + r] = va[it] * dinv_s[k];
      Xb[k * kSdLd + r] = vb[it];
    }
    __syncthreads();
    if (l0 + kSdSlab < mg) fetch(l0 + kSdSlab);
#pragma unroll
    for (int ks = 0; ks < kSdSlab / 4; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) af[a] = Xa[(4 * ks + fk) * kSdLd + wi + 8 * a + fr];
#pragma unroll
      for (int c = 0; c < 4; ++c) bf[c] = Xb[(4 * ks + fk) * kSdLd + wj + 8 * c + fr];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) dmma884(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
    }
    if (ti == tj && tid < kSdTile) {   // b[R] -= X D^-1 b_diag for the rows of this (diagonal) tile
#pragma unroll
      for (int k = 0; k < kSdSlab; ++k) bacc = fma(Xa[k * kSdLd + tid], bd_s[k], bacc);
    }
    __syncthreads();
  }
  // epilogue: D fragment = (row lane / 4, cols 2 (lane % 4) + {0, 1}) of every 8 x 8 tile
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = kSdTile * ti + wi + 8 * a + fr, j = kSdTile * tj + wj + 8 * c + 2 * fk + e;
        if (i < nr && j < nr) {
          A[R(i) + (size_t)pos * R(j)] -= acc[a][c][e];
          if (ti != tj) A[R(j) + (size_t)pos * R(i)] -= acc[a][c][e];
        }
      }
  if (ti == tj && tid < kSdTile) {
    const int i = kSdTile * ti + tid;
    if (i < nr) b[R(i)] -= bacc;
  }
  (void)m;
}

// -------------------------------------------------------------------------------------------------
constexpr int kMgThreads = 256;
constexpr int kMgMaxDense = 32;

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// smem (doubles): P[32*32] V[32*32] cs[6*16] (32 spare) wk[2*32*32]
__global__ void __launch_bounds__(kMgThreads)
marg_schur_eig_kernel(isv_marg_generic_in in, isv_marg_generic_out out, double* __restrict__ Gbuf, int schur_only,
                      int phase_a_done) {   // Gbuf: n x md scratch per problem (stride n * max(n, md))
  extern __shared__ double smem[];
  double* P = smem;                       // m_dense x m_dense
  double* V = P + kMgMaxDense * kMgMaxDense;
  double* cs = V + kMgMaxDense * kMgMaxDense;
  double* wk = cs + 6 * 16 + 32;          // 2 x m_dense x m_dense: Gauss-Jordan work space
  __shared__ int s_status;
  if (threadIdx.x == 0) s_status = 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = kMgThreads / 32;
  const int prob = blockIdx.x;
  const int pos = in.pos, md = in.m_dense, mg = in.m_diag, m = md + mg, n = pos - m, nr = md + n;
  double* A = out.A + (size_t)prob * pos * pos;
  double* b = out.b + (size_t)prob * pos;
  double* Ar = out.A_red + (size_t)prob * n * n;
  double* br = out.b_red + (size_t)prob * n;
  double* G = Gbuf + (size_t)prob * n * (n > md ? n : md);   // T = A_rm pinv (n x md); the stride covers md > n too
  const double eps = in.eps;
  int status = 0;
  auto R = [&](int i) { return i < md ? i : i + mg; };   // compact index -> position in A

  // ---- phase A: eliminate the diagonal (scalar) marginalized blocks:  C -= X D^-1 X^T  on DMMA ----
  if (mg > 0 && !phase_a_done) {   // in-kernel variant (one CTA per problem, fragments straight from L2)
    const int nt = (nr + 7) / 8;
    const int fr = lane >> 2, fk = lane & 3;
    for (int t = warp; t < nt * nt; t += nw) {
      const int ti = t / nt, tj = t - ti * nt;
      if (tj < ti) continue;               // symmetric: upper tiles, mirrored below
      const int ra = 8 * ti + fr, rb = 8 * tj + fr;
      const double* rowa = (ra < nr) ? A + R(ra) : nullptr;   // A[R(ra) + pos * l]
      const double* rowb = (rb < nr) ? A + R(rb) : nullptr;
      double d0 = 0.0, d1 = 0.0;
      for (int k0 = 0; k0 < mg; k0 += 4) {
        const int l = k0 + fk;
        double a = 0.0, bb = 0.0;
        if (l < mg) {
          const size_t col = (size_t)pos * (md + l);
          const double d = A[(md + l) + col];
          const double di = d > eps ? 1.0 / d : 0.0;
          if (rowa) a = rowa[col] * di;
          if (rowb) bb = rowb[col];
        }
        dmma884(d0, d1, a, bb);
      }
      // D fragment: row = lane / 4, cols = 2 (lane % 4) + {0, 1}
      const int i = 8 * ti + fr, j0 = 8 * tj + 2 * fk;
      // tiles are disjoint in (i, j): the entries written here are only read as INPUT in columns
      // md..m-1, which are never written, so the in-place update is race-free
      if (i < nr && j0 < nr) {
        A[R(i) + (size_t)pos * R(j0)] -= d0;
        if (ti != tj) A[R(j0) + (size_t)pos * R(i)] -= d0;
      }
      if (i < nr && j0 + 1 < nr) {
        A[R(i) + (size_t)pos * R(j0 + 1)] -= d1;
        if (ti != tj) A[R(j0 + 1) + (size_t)pos * R(i)] -= d1;
      }
    }
    // b_R -= X D^-1 b_diag ; flag unusable pivots
    for (int i = tid; i < nr; i += kMgThreads) {
      double acc = 0.0;
      for (int l = 0; l < mg; ++l) {
        const size_t col = (size_t)pos * (md + l);
        const double d = A[(md + l) + col];
        if (d > eps) acc = fma(A[R(i) + col] / d, b[md + l], acc);
      }
      b[R(i)] -= acc;
    }
    for (int l = tid; l < mg; l += kMgThreads)
      if (!(A[(md + l) + (size_t)pos * (md + l)] > eps)) status |= ISV_W_RANK_DEFICIENT;
  }
  __syncthreads();
  // ---- phase B: dense marginalized block: Amm = (P + P^T)/2, eigen-thresholded pseudo-inverse ---------
  for (int idx = tid; idx < md * md; idx += kMgThreads) {
    const int i = idx % md, j = idx / md;
    P[i + md * j] = 0.5 * (A[i + (size_t)pos * j] + A[j + (size_t)pos * i]);
  }
  __syncthreads();
  if (warp == 0 && md > 0) {
    // fast path: A_mm^-1 by Gauss-Jordan, accepted when the smallest eigenvalue is provably > eps (see below), so
    // VINS-Mono's eigen-thresholded pseudo-inverse keeps every eigenvalue and IS the inverse.
    // Otherwise (rank-deficient, indefinite or badly scaled block) the literal eigen-decomposition below decides.
    for (int idx = lane; idx < md * md; idx += 32) V[idx] = P[idx];
    __syncwarp();
    const int sing = w_inverse(V, md, md, wk, lane);
    double fn = 0.0, an = 0.0;
    for (int idx = lane; idx < md * md; idx += 32) { fn = fma(V[idx], V[idx], fn); an = fma(P[idx], P[idx], an); }
    fn = warp_sum(fn);
    an = warp_sum(an);
    // |lam|_min >= 1 / ||A^-1||_F must exceed eps, and cond_F < 1e12 so that rounding (<= n eps_mach ||A||) cannot have
    // flipped the sign of an eigenvalue of this sum of J^T J: then every eigenvalue is positive and > eps
    const bool fast = !sing && isfinite(fn) && fn > 0.0 && fn * eps * eps < 1.0 && fn * an < 1e24;
    if (fast) {
      __syncwarp();
      for (int idx = lane; idx < md * md; idx += 32) {
        const int i = idx % md, j = idx / md;
        P[idx] = 0.5 * (V[i + md * j] + V[j + md * i]);
      }
    } else {
    if (w_jacobi_eig(P, md, V, md, md, cs, lane) >= 30) status |= ISV_W_EIG_NOCONV;
    // pinv = V diag(lam > eps ? 1/lam : 0) V^T  -> back into P (cs reused for the inverted spectrum)
    for (int k = lane; k < md; k += 32) {
      const double lam = P[k + md * k];
      if (!(lam > eps)) status |= ISV_W_RANK_DEFICIENT;
      cs[k] = lam > eps ? 1.0 / lam : 0.0;
    }
    __syncwarp();
    for (int idx = lane; idx < md * md; idx += 32) {
      const int i = idx % md, j = idx / md;
      double acc = 0.0;
      for (int k = 0; k < md; ++k) acc = fma(V[i + md * k] * cs[k], V[j + md * k], acc);
      P[idx] = acc;
    }
    }
  }
  __syncthreads();
  // A_red = A_rr - A_rm pinv A_mr ;  b_red = b_rr - A_rm pinv b_mm      (r = kept, m = dense marginalized)
  for (int idx = tid; idx < n * md; idx += kMgThreads) {   // T = A_rm pinv  (n x md) into G scratch
    const int i = idx % n, k = idx / n;
    double acc = 0.0;
    for (int l = 0; l < md; ++l) acc = fma(A[(m + i) + (size_t)pos * l], P[l + md * k], acc);
    G[idx] = acc;
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += kMgThreads) {
    const int i = idx % n, j = idx / n;
    double acc = A[(m + i) + (size_t)pos * (m + j)];
    for (int k = 0; k < md; ++k) acc = fma(-G[i + n * k], A[k + (size_t)pos * (m + j)], acc);
    Ar[idx] = acc;
  }
  for (int i = tid; i < n; i += kMgThreads) {
    double acc = b[m + i];
    for (int k = 0; k < md; ++k) acc = fma(-G[i + n * k], b[k], acc);
    br[i] = acc;
  }
  __syncthreads();
  // the eigen-decomposition of the reduced system is its own kernel (isv_sym_eig.cuh); schur_only callers (the
  // reduced camera system of DENSE_SCHUR) get rank = -1
  if (schur_only && tid == 0) out.rank[prob] = -1;
  if (status) atomicOr(&s_status, status);
  __syncthreads();
  if (tid == 0 && out.status && s_status) atomicOr(out.status + prob, s_status);
  (void)br; (void)nw;
}

// -------------------------------------------------------------------------------------------------
// MarginalizationFactor (VINS-Mono marginalization_factor.cpp, `MarginalizationFactor::Evaluate`): the
// previous prior |r0 + J dx|^2 as a residual block of the next round.
// marg_prior_eval_kernel : dx (every CTA recomputes the few hundred entries in shared memory), then
//                          residuals = r0 + J dx (thread per row: coalesced down the column-major J) and
//                          the per-block ceres Jacobians (a re-layout of J, last pose column zero).
// marg_prior_add_kernel  : A[pos, pos] += J^T J, b[pos] += J^T r -- 32 x 32 tiles of column pairs (upper
//                          triangle, mirrored), rows staged through shared memory in slabs of 32.
// -------------------------------------------------------------------------------------------------
constexpr int kPriorThreads = 256;

__device__ __forceinline__ int prior_local(int global_size) { return global_size == 7 ? 6 : global_size; }

__global__ void __launch_bounds__(kPriorThreads)
marg_prior_eval_kernel(isv_marg_prior pr, double* __restrict__ residuals, double* __restrict__ jacobians,
                       int32_t* status) {
  extern __shared__ double dx[];   // n doubles, then n_blocks + 1 ints (Jacobian record offsets)
  const int n = pr.n, tid = threadIdx.x;
  int* joff = reinterpret_cast<int*>(dx + n);
  for (int i = tid; i < n; i += blockDim.x) dx[i] = 0.0;
  if (tid == 0) {
    int o = 0;
    for (int b = 0; b < pr.n_blocks; ++b) { joff[b] = o; o += n * pr.blocks[b].global_size; }
    joff[pr.n_blocks] = o;
  }
  __syncthreads();
  for (int b = tid; b < pr.n_blocks; b += blockDim.x) {
    const isv_prior_block bl = pr.blocks[b];
    const int ls = prior_local(bl.global_size);
    if (bl.global_size < 1 || bl.idx < 0 || bl.idx + ls > n || bl.x_offset < 0) {
      if (status) atomicOr(status, ISV_W_BAD_INDEX);
      continue;
    }
    const double* x = pr.x + bl.x_offset;
    const double* x0 = pr.x0 + bl.x_offset;
    if (bl.global_size == 7) {
      for (int c = 0; c < 3; ++c) dx[bl.idx + c] = x[c] - x0[c];
      // Eigen: q0.inverse() = conj / |q0|^2 (no normalisation)
      const Quat d = qmul(qinv(quat_from_pose(x0)), quat_from_pose(x));
      const double sgn = (d.w >= 0.0) ? 2.0 : -2.0;
      dx[bl.idx + 3] = sgn * d.x; dx[bl.idx + 4] = sgn * d.y; dx[bl.idx + 5] = sgn * d.z;
    } else {
      for (int c = 0; c < ls; ++c) dx[bl.idx + c] = x[c] - x0[c];
    }
  }
  __syncthreads();
  const double* J = pr.linearized_jacobians;
  for (int i = blockIdx.x * blockDim.x + tid; i < n; i += gridDim.x * blockDim.x) {
    double a0 = 0.0, a1 = 0.0;
    int j = 0;
    for (; j + 1 < n; j += 2) {
      a0 = fma(J[i + (size_t)n * j], dx[j], a0);
      a1 = fma(J[i + (size_t)n * (j + 1)], dx[j + 1], a1);
    }
    if (j < n) a0 = fma(J[i + (size_t)n * j], dx[j], a0);
    residuals[i] = pr.linearized_residuals[i] + (a0 + a1);
  }
  if (jacobians) {
    const int total = joff[pr.n_blocks];
    for (int e = blockIdx.x * blockDim.x + tid; e < total; e += gridDim.x * blockDim.x) {
      int b = 0;
      while (e >= joff[b + 1]) ++b;
      const isv_prior_block bl = pr.blocks[b];
      const int gs = bl.global_size, ls = prior_local(gs);
      const int rem = e - joff[b], row = rem / gs, col = rem - row * gs;
      const bool ok = col < ls && bl.idx >= 0 && bl.idx + ls <= n;
      jacobians[e] = ok ? J[row + (size_t)n * (bl.idx + col)] : 0.0;
    }
  }
}

constexpr int kPaTile = 32;

__global__ void __launch_bounds__(kPriorThreads)
marg_prior_add_kernel(isv_marg_prior pr, const double* __restrict__ res, double* __restrict__ A, double* __restrict__ b,
                      int pos, int32_t* status, int dlo, int dhi) {   // [dlo, dhi): the diagonal marginalized block
  // blockIdx.z: the problem of the batch this CTA adds the (shared) prior to
  A += (size_t)blockIdx.z * pos * pos;
  b += (size_t)blockIdx.z * pos;
  if (status) status += blockIdx.z;
  __shared__ double Ja[kPaTile][kPaTile + 1], Jb[kPaTile][kPaTile + 1], rs[kPaTile];
  __shared__ int cpa[kPaTile], cpb[kPaTile];
  const int ci0 = blockIdx.x * kPaTile, cj0 = blockIdx.y * kPaTile;
  if (ci0 > cj0) return;
  const int n = pr.n, tid = threadIdx.x;
  if (tid < 2 * kPaTile) {
    const int col = (tid < kPaTile) ? ci0 + tid : cj0 + tid - kPaTile;
    int p = -1;
    for (int k = 0; k < pr.n_blocks; ++k) {
      const isv_prior_block bl = pr.blocks[k];
      const int ls = prior_local(bl.global_size);
      if (col >= bl.idx && col < bl.idx + ls && bl.pos >= 0) {
        if (bl.pos + ls > pos) {
          if (status) atomicOr(status, ISV_W_BAD_INDEX);
        } else {
          p = bl.pos + col - bl.idx;
          // the prior is a dense J^T J over all of its blocks: a block of it inside the diagonal marginalized range is
          // coupled with everything else the prior holds
          if (p >= dlo && p < dhi && status) atomicOr(status, ISV_W_DIAG_COUPLED);
        }
      }
    }
    if (tid < kPaTile) cpa[tid] = p; else cpb[tid - kPaTile] = p;
  }
  const int a = tid & 31, cq = tid >> 5;   // this thread: column a of tile i, columns cq + 8 q of tile j
  double acc[4] = {0.0, 0.0, 0.0, 0.0}, accb = 0.0;
  const double* J = pr.linearized_jacobians;
  for (int k0 = 0; k0 < n; k0 += kPaTile) {
    __syncthreads();
    const int kk = tid & 31, row = k0 + kk;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = cq + 8 * q;
      Ja[kk][c] = (row < n && ci0 + c < n) ? J[row + (size_t)n * (ci0 + c)] : 0.0;
      Jb[kk][c] = (row < n && cj0 + c < n) ? J[row + (size_t)n * (cj0 + c)] : 0.0;
    }
    if (tid < kPaTile) rs[tid] = (k0 + tid < n) ? res[k0 + tid] : 0.0;
    __syncthreads();
#pragma unroll 8
    for (int l = 0; l < kPaTile; ++l) {
      const double va = Ja[l][a];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fma(va, Jb[l][cq + 8 * q], acc[q]);
      if (cq == 0) accb = fma(va, rs[l], accb);
    }
  }
  const int pa = cpa[a];
  if (pa < 0) return;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int pc = cpb[cq + 8 * q];
    if (pc < 0) continue;
    A[pa + (size_t)pos * pc] += acc[q];
    if (ci0 != cj0) A[pc + (size_t)pos * pa] += acc[q];
  }
  if (ci0 == cj0 && cq == 0) b[pa] += accb;
}

}  // namespace isv
