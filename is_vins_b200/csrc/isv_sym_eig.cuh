// Symmetric eigen-decomposition of the reduced system of the generic marginalization engine -- what VINS-Mono's
// `Eigen::SelfAdjointEigenSolver<MatrixXd> saes2(A)` does in MarginalizationInfo::marginalize (published
// marginalization_factor.cpp; IS-VINS keeps the same solver for its own truncations, src/estimator.cpp:920,1311,
// 1479) -- with the same algorithm class as Eigen: Householder tridiagonalization, implicit QL sweeps on the
// tridiagonal, rotations accumulated into the orthogonal factor.  Split by what bounds each part:
//
//   sym_tridiag_kernel   one CTA per problem.  W = Q T Q^T (LAPACK dsytd2, full symmetric storage in L2), thread per
//                        row (coalesced across threads, 8 independent loads in flight per thread, no shuffles); the
//                        rank-2 update of step k-1 is fused into the symv of step k, so the trailing block is read
//                        once and written once per step -- L2-bandwidth bound.
//   tridiag_ql_kernel    one WARP per problem.  The implicit QL recurrence on (d, e) is a serial dependent chain
//                        (~210 cycles per rotation: rsqrt + 8 DFMA), so it gets the smallest possible footprint --
//                        thousands of problems run concurrently -- and does not touch Z at all: lane 0 walks the
//                        chain and LOGS every rotation (c, s, column); the other lanes help with the deflation scan.
//   q_rows_kernel /      grid (row slabs, problems).  Rows of Z = Q Y are independent: a CTA builds a slab of rows of Q
//   ql_apply_kernel      in shared memory (e_r^T through the n-1 reflectors; on a side stream, hidden under the QL
//                        kernel), then every thread streams the whole rotation log through its own row (one LDS + one
//                        STS + 4 DFMA per rotation, the running column in a register, one DFMA on the dependent chain).
//   eig_prior_kernel     outputs in ascending eigenvalue order: linearized_jacobians = S^1/2 V^T,
//                        linearized_residuals = S^-1/2 V^T b, S thresholded at eps (dropped rows exactly zero).
//
// History: pivoted Cholesky + CTA-level one-sided Jacobi (~30 n^3 flop, n-1 barrier-separated rounds per sweep):
// 315 ms per 148 problems at n = 307 (BASELINE configs[3] variant b); one CTA doing tridiagonalization + QL with the
// rotations applied to Z in L2 after every sweep: 46 ms.
#pragma once
#include "isv_device_math.cuh"

#include "../../include/isv_capi.h"

namespace isv {

constexpr int kSeMaxN = 1024;
constexpr int kSeThreads = 512;
constexpr int kSeLogFactor = 3;      // rotation log capacity = kSeLogFactor * n^2 + 64 (observed ~0.9 n^2)
constexpr int kSeSweepFactor = 8;    // sweep table capacity = kSeSweepFactor * n + 8 (observed ~1.7 n)

__host__ __device__ inline size_t sym_eig_log_cap(int n) { return (size_t)kSeLogFactor * n * n + 64; }
__host__ __device__ inline size_t sym_eig_sweep_cap(int n) { return (size_t)kSeSweepFactor * n + 8; }
// per problem scratch (doubles): d[n] e[n] tau[n] | log (c, s)[cap] | then ints: sweep table int4[kSeSweepFactor n], meta[4]
__host__ __device__ inline size_t sym_eig_scratch_bytes(int n) {
  const size_t cap = sym_eig_log_cap(n);
  return (((size_t)3 * n + (n & 1) + 2 * cap) * sizeof(double) + (sym_eig_sweep_cap(n) * 4 + 4) * sizeof(int) + 255) & ~(size_t)255;
}
struct SymEigScratch {
  double* d; double* e; double* tau; double2* cs;
  int4* sweep;   // per QL sweep: x = first log entry, y = m (rotations on column pairs (m-1, m), (m-2, m-1), ...), z = count
  int* meta;     // meta[0] = rotations logged, meta[1] = fail flags, meta[2] = sweeps
};
__host__ __device__ inline SymEigScratch sym_eig_scratch(char* base, int n, size_t prob) {
  const size_t cap = sym_eig_log_cap(n);
  char* p = base + prob * sym_eig_scratch_bytes(n);
  SymEigScratch s;
  s.d = reinterpret_cast<double*>(p);
  s.e = s.d + n;
  s.tau = s.e + n;
  s.cs = reinterpret_cast<double2*>(s.tau + n + (n & 1));   // 16-byte aligned
  s.sweep = reinterpret_cast<int4*>(s.cs + cap);
  s.meta = reinterpret_cast<int*>(s.sweep + sym_eig_sweep_cap(n));
  return s;
}

__device__ __forceinline__ double se_block_sum(double v, double* red, int nw) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}

// smem doubles: V[2][n] Wv[2][n] col[n] red[32] part[kSeThreads]
__host__ __device__ inline size_t sym_tridiag_smem_doubles(int n) { return (size_t)5 * n + 32 + kSeThreads; }

// Householder tridiagonalization W = Q T Q^T (LAPACK dsytd2, lower variant) with the rank-2 update of step k-1 FUSED
// into the symv of step k: one read + one write pass over the trailing block per step instead of two reads + one
// write.  W: n x n column-major symmetric, full storage (destroyed: column k below the diagonal ends up holding the
// reflector v_k with v_k[0] = 1 stored explicitly); d, e, tau: the tridiagonal and the reflector scalars.
__device__ void sym_tridiag_cta(double* __restrict__ W, int n, double* __restrict__ d, double* __restrict__ e,
                                double* __restrict__ tau, double* smem) {
  const int tid = threadIdx.x, nt = blockDim.x, nw = nt >> 5;
  double* Vb = smem;             // [2][n]  pending / new reflector, by absolute row index
  double* Wb = Vb + 2 * n;       // [2][n]  pending / new w
  double* col = Wb + 2 * n;      // [n]     column k of the up-to-date trailing matrix
  double* red = col + n;         // 32
  double* part = red + 32;       // nt
  bool pend = false;             // (Vb[cur], Wb[cur]) still has to be subtracted from W[k:, k:]
  int cur = 0;
  for (int k = 0; k + 1 < n; ++k) {
    const int len = n - k - 1;
    const double* Vp = Vb + cur * n;
    const double* Wp = Wb + cur * n;
    double* Vn = Vb + (cur ^ 1) * n;
    double* Wn = Wb + (cur ^ 1) * n;
    // (a) column k of the up-to-date matrix, rows k..n-1
    for (int i = k + tid; i < n; i += nt) {
      double a = W[(size_t)i + (size_t)n * k];
      if (pend) a -= fma(Vp[i], Wp[k], Wp[i] * Vp[k]);
      col[i] = a;
    }
    __syncthreads();
    // (b) the reflector of x = col[k+1:]
    double acc = 0.0;
    for (int i = k + 2 + tid; i < n; i += nt) acc = fma(col[i], col[i], acc);
    const double xn2 = se_block_sum(acc, red, nw);
    const double alpha = col[k + 1];
    double beta = alpha, tk = 0.0, scale = 0.0;
    if (xn2 > 0.0) {
      beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
      tk = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    for (int i = k + 1 + tid; i < n; i += nt) {
      const double v = (i == k + 1) ? 1.0 : col[i] * scale;
      Vn[i] = v;
      W[(size_t)i + (size_t)n * k] = v;                      // kept for the Q rows (ql_apply_kernel)
    }
    if (tid == 0) { d[k] = col[k]; e[k] = beta; tau[k] = tk; }
    __syncthreads();
    // (c) trailing block rows / cols k+1..n-1: apply the pending update, accumulate (A22 v) per row.
    //     thread (row i, part q): rows across consecutive threads (coalesced), the columns j = q (mod P) per part
    if (pend || tk != 0.0) {
      double* A22 = W + (size_t)(k + 1) + (size_t)n * (k + 1);
      const double* vp = Vp + k + 1; const double* wp = Wp + k + 1; const double* vn = Vn + k + 1;
      const int R = (len + 31) & ~31;
      const bool fits = R <= nt;
      const int P = fits ? min(nt / R, 8) : 1;
      const int i0 = fits ? tid % R : tid, q = fits ? tid / R : 0, istep = fits ? (1 << 30) : nt;
      if (q < P)
        for (int i = i0; i < len; i += istep) {
          double* row = A22 + i;
          const double vi = pend ? vp[i] : 0.0, wi = pend ? wp[i] : 0.0;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          int j = q;
          for (; j + 7 * P < len; j += 8 * P) {
            double u[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) u[t] = row[(size_t)n * (j + t * P)];
            if (pend) {
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                u[t] -= fma(vi, wp[j + t * P], wi * vp[j + t * P]);
                row[(size_t)n * (j + t * P)] = u[t];
              }
            }
            a0 = fma(u[0], vn[j], a0);         a1 = fma(u[1], vn[j + P], a1);
            a2 = fma(u[2], vn[j + 2 * P], a2); a3 = fma(u[3], vn[j + 3 * P], a3);
            a0 = fma(u[4], vn[j + 4 * P], a0); a1 = fma(u[5], vn[j + 5 * P], a1);
            a2 = fma(u[6], vn[j + 6 * P], a2); a3 = fma(u[7], vn[j + 7 * P], a3);
          }
          for (; j < len; j += P) {
            double u = row[(size_t)n * j];
            if (pend) { u -= fma(vi, wp[j], wi * vp[j]); row[(size_t)n * j] = u; }
            a0 = fma(u, vn[j], a0);
          }
          const double s = (a0 + a1) + (a2 + a3);
          if (q > 0) part[(q - 1) * R + i] = s;   // (P - 1) * R <= nt
          else col[k + 1 + i] = s;                // col[] is free again: A22 v
        }
      __syncthreads();
      // (d) w = p - (tau / 2)(p . v) v ,  p = tau A22 v
      double pv = 0.0;
      if (tk != 0.0)
        for (int i = tid; i < len; i += nt) {
          double s = col[k + 1 + i];
          for (int qq = 1; qq < P; ++qq) s += part[(qq - 1) * R + i];
          s *= tk;
          col[k + 1 + i] = s;
          pv = fma(s, vn[i], pv);
        }
      pv = se_block_sum(pv, red, nw);
      const double c2 = -0.5 * tk * pv;
      for (int i = tid; i < len; i += nt) Wn[k + 1 + i] = fma(c2, vn[i], col[k + 1 + i]);
    }
    pend = tk != 0.0;
    cur ^= 1;
    __syncthreads();
  }
  if (tid == 0) {
    const double* Vp = Vb + cur * n;
    const double* Wp = Wb + cur * n;
    d[n - 1] = W[(size_t)(n - 1) + (size_t)n * (n - 1)] - (pend ? 2.0 * Vp[n - 1] * Wp[n - 1] : 0.0);
    e[n - 1] = 0.0;
    tau[n - 1] = 0.0;
  }
}

// A_red (read-only) -> W = (A_red + A_red^T) / 2 -> tridiagonal (d, e, tau in the scratch), reflectors left in W
__global__ void __launch_bounds__(kSeThreads)
sym_tridiag_kernel(int n, const double* __restrict__ A_all, size_t a_stride, double* __restrict__ Wbuf, size_t w_stride,
                   char* __restrict__ scratch) {
  extern __shared__ double smem[];
  const int prob = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const double* Ar = A_all + (size_t)prob * a_stride;
  double* W = Wbuf + (size_t)prob * w_stride;
  const SymEigScratch sc = sym_eig_scratch(scratch, n, prob);
  // Eigen reads the lower triangle only; VINS-Mono symmetrises Amm the same way
  for (size_t idx = tid; idx < (size_t)n * n; idx += nt) {
    const int i = (int)(idx % n), j = (int)(idx / n);
    W[idx] = 0.5 * (Ar[idx] + Ar[j + (size_t)n * i]);
  }
  __syncthreads();
  sym_tridiag_cta(W, n, sc.d, sc.e, sc.tau, smem);
}

// implicit QL with Wilkinson shift on (d, e) (the tqli recurrence; deflation test relative to the neighbours as in
// LAPACK dsteqr); every rotation is logged: cs[r] = (c, s), and every sweep gets a row in the sweep table.  One warp
// per problem; smem: d[n] e[n].

__global__ void __launch_bounds__(32)
tridiag_ql_kernel(int n, char* __restrict__ scratch) {
  extern __shared__ double smem[];
  const int prob = blockIdx.x, lane = threadIdx.x;
  const SymEigScratch sc = sym_eig_scratch(scratch, n, prob);
  double* d = smem;
  double* e = smem + n;
  for (int i = lane; i < n; i += 32) { d[i] = sc.d[i]; e[i] = sc.e[i]; }
  __syncwarp();
  const double epsm = 2.220446049250313e-16;
  const long long cap = (long long)sym_eig_log_cap(n);
  const int sweep_cap = (int)sym_eig_sweep_cap(n);
  long long count = 0;
  int fail = 0, nsweep = 0;
  for (int l = 0; l < n && !(fail & 2); ++l) {
    for (int iter = 0;; ++iter) {
      int m = n - 1;
      for (int base = l; base < n - 1; base += 32) {
        const int mm = base + lane;
        const bool small = mm < n - 1 && fabs(e[mm]) <= epsm * (fabs(d[mm]) + fabs(d[mm + 1]));
        const unsigned hit = __ballot_sync(kFullMask, small);
        if (hit) { m = base + __ffs(hit) - 1; break; }
      }
      if (m == l) break;
      if (iter >= 60) { fail |= 1; break; }                 // give this eigenvalue up (flagged)
      if (count + (m - l) > cap || nsweep >= sweep_cap) { fail |= 2; break; }   // log full: stop (flagged); not a decomposition
      int cnt = 0;
      if (lane == 0) {
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = sqrt(fma(g, g, 1.0));
        g = d[m] - d[l] + e[l] / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        bool broke = false;
        double2* cs = sc.cs + count;
        double ei = e[m - 1], di = d[m - 1], di1 = d[m];
        for (int i = m - 1; i >= l; --i) {
          const double f = s * ei, b = c * ei;
          const double en = i > l ? e[i - 1] : 0.0, dn = i > l ? d[i - 1] : 0.0;   // next iteration's operands, off the chain
          const double r2 = fma(f, f, g * g);
          if (r2 == 0.0) { d[i + 1] = di1 - p; e[m] = 0.0; broke = true; break; }
          const double y = rsqrt(r2);
          e[i + 1] = r2 * y;
          s = f * y;
          c = g * y;
          g = di1 - p;
          r = fma(di - g, s, 2.0 * c * b);
          p = s * r;
          di1 = g + p;
          d[i + 1] = di1;
          g = fma(c, r, -b);
          cs[cnt] = make_double2(c, s);
          ++cnt;
          ei = en; di1 = di; di = dn;
        }
        if (!broke) { d[l] -= p; e[l] = g; e[m] = 0.0; }
        sc.sweep[nsweep] = make_int4((int)count, m, cnt, 0);
      }
      cnt = __shfl_sync(kFullMask, cnt, 0);
      count += cnt;
      ++nsweep;
      __syncwarp();
    }
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) sc.d[i] = d[i];      // eigenvalues (unsorted)
  if (lane == 0) { sc.meta[0] = (int)count; sc.meta[1] = fail; sc.meta[2] = nsweep; }
}

// The eigenvector matrix Z = Q Y by row slabs held in shared memory (rows are independent under both factors):
//   q_rows_kernel     rows of Q = H_0 H_1 ... H_{n-2}: e_r^T through every reflector (v_k staged through a double buffer).
//                     Needs only the tridiagonalization, so it runs on a side stream UNDER the serial QL kernel.
//   ql_apply_kernel   the whole rotation log Y streamed through the rows, sweep by sweep ((c, s) of the next sweep
//                     arrives by cp.async while the current one is applied).
// Dynamic smem of both: rows x (n | 1) doubles (odd stride: the threads' rows fall into distinct banks).
constexpr int kQaThreads = 64;
constexpr int kQaPre = (kSeMaxN + kQaThreads - 1) / kQaThreads;   // reflector elements prefetched per thread
__host__ __device__ inline int ql_apply_rows(int n) {
  const size_t budget = 190 * 1024;    // + 32 KB static (reflector / rotation double buffers) < 227 KB
  const int rows = (int)(budget / (sizeof(double) * (size_t)(n | 1)));
  return rows >= 64 ? 64 : (rows >= 32 ? 32 : rows);
}
__host__ __device__ inline size_t ql_apply_smem_bytes(int n, int rows) { return sizeof(double) * (size_t)rows * (n | 1); }

__device__ __forceinline__ void slab_store(const double* slab, int ld, double* __restrict__ Z, int n, int row0, int nr) {
  // for a fixed column the rows are contiguous in the column-major Z
  for (int idx = threadIdx.x; idx < nr * n; idx += blockDim.x) {
    const int r = idx % nr, c = idx / nr;
    Z[(size_t)(row0 + r) + (size_t)n * c] = slab[(size_t)r * ld + c];
  }
}

__global__ void __launch_bounds__(kQaThreads)
q_rows_kernel(int n, int rows, const double* __restrict__ Wbuf, size_t w_stride, double* __restrict__ Zbuf,
              const char* __restrict__ scratch) {
  extern __shared__ double smem[];
  __shared__ double vb[2 * kSeMaxN];     // static: provably distinct from the slab, so its reads hoist above the slab stores
  const int prob = blockIdx.y, row0 = blockIdx.x * rows, tid = threadIdx.x, nt = blockDim.x;
  const int ld = n | 1;
  double* slab = smem;
  const double* W = Wbuf + (size_t)prob * w_stride;
  const SymEigScratch sc = sym_eig_scratch(const_cast<char*>(scratch), n, prob);
  const int nr = min(rows, n - row0);
  const bool live = tid < nr;
  double* mine = slab + (size_t)tid * ld;
  if (tid < rows)
    for (int c = 0; c < n; ++c) mine[c] = (c == row0 + tid) ? 1.0 : 0.0;
  // x <- x - tau_k (x . v_k) v_k on the entries k+1.. , k = 0 .. n-2
  if (n > 1)
    for (int j = tid; j < n - 1; j += nt) vb[j] = W[(size_t)(1 + j)];          // v_0 = W[1:, 0]
  __syncthreads();
  for (int k = 0; k + 1 < n; ++k) {
    const int len = n - k - 1;
    const double* v = vb + (k & 1) * kSeMaxN;
    double pre[kQaPre];
    const bool more = k + 2 < n;
    if (more) {                                            // v_{k+1} in flight while step k computes
      const double* src = W + (size_t)(k + 2) + (size_t)n * (k + 1);
#pragma unroll
      for (int t = 0; t < kQaPre; ++t) {
        const int j = tid + t * nt;
        pre[t] = j < len - 1 ? src[j] : 0.0;
      }
    }
    const double tk = sc.tau[k];
    if (live && tk != 0.0) {
      double* x = mine + k + 1;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int j = 0;
      for (; j + 7 < len; j += 8) {
        const double x0 = x[j], x1 = x[j + 1], x2 = x[j + 2], x3 = x[j + 3], x4 = x[j + 4], x5 = x[j + 5], x6 = x[j + 6], x7 = x[j + 7];
        const double2 v01 = *reinterpret_cast<const double2*>(v + j), v23 = *reinterpret_cast<const double2*>(v + j + 2),
                      v45 = *reinterpret_cast<const double2*>(v + j + 4), v67 = *reinterpret_cast<const double2*>(v + j + 6);
        a0 = fma(x0, v01.x, a0); a1 = fma(x1, v01.y, a1); a2 = fma(x2, v23.x, a2); a3 = fma(x3, v23.y, a3);
        a0 = fma(x4, v45.x, a0); a1 = fma(x5, v45.y, a1); a2 = fma(x6, v67.x, a2); a3 = fma(x7, v67.y, a3);
      }
      for (; j < len; ++j) a0 = fma(x[j], v[j], a0);
      const double t = -tk * ((a0 + a1) + (a2 + a3));
      for (j = 0; j + 7 < len; j += 8) {
        const double x0 = x[j], x1 = x[j + 1], x2 = x[j + 2], x3 = x[j + 3], x4 = x[j + 4], x5 = x[j + 5], x6 = x[j + 6], x7 = x[j + 7];
        const double2 v01 = *reinterpret_cast<const double2*>(v + j), v23 = *reinterpret_cast<const double2*>(v + j + 2),
                      v45 = *reinterpret_cast<const double2*>(v + j + 4), v67 = *reinterpret_cast<const double2*>(v + j + 6);
        x[j] = fma(t, v01.x, x0); x[j + 1] = fma(t, v01.y, x1); x[j + 2] = fma(t, v23.x, x2); x[j + 3] = fma(t, v23.y, x3);
        x[j + 4] = fma(t, v45.x, x4); x[j + 5] = fma(t, v45.y, x5); x[j + 6] = fma(t, v67.x, x6); x[j + 7] = fma(t, v67.y, x7);
      }
      for (; j < len; ++j) x[j] = fma(t, v[j], x[j]);
    }
    if (more) {
      double* dst = vb + ((k + 1) & 1) * kSeMaxN;
#pragma unroll
      for (int t = 0; t < kQaPre; ++t) {
        const int j = tid + t * nt;
        if (j < len - 1) dst[j] = pre[t];
      }
    }
    __syncthreads();
  }
  slab_store(slab, ld, Zbuf + (size_t)prob * n * n, n, row0, nr);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}

__global__ void __launch_bounds__(kQaThreads)
ql_apply_kernel(int n, int rows, double* __restrict__ Zbuf, const char* __restrict__ scratch) {
  extern __shared__ double smem[];
  __shared__ double2 ccs[2][kSeMaxN];    // (c, s) of the current / next sweep; static for the same reason as vb
  const int prob = blockIdx.y, row0 = blockIdx.x * rows, tid = threadIdx.x, nt = blockDim.x;
  const int ld = n | 1;
  double* slab = smem;
  double* Z = Zbuf + (size_t)prob * n * n;
  const SymEigScratch sc = sym_eig_scratch(const_cast<char*>(scratch), n, prob);
  const int nr = min(rows, n - row0);
  const bool live = tid < nr;
  double* mine = slab + (size_t)tid * ld;
  const int nsweep = sc.meta[2];
  // sweep headers travel two iterations ahead in registers: neither the staging nor the compute waits on a global load
  const int4 none = make_int4(0, 0, 0, 0);
  int4 h_cur = nsweep > 0 ? sc.sweep[0] : none, h_nxt = nsweep > 1 ? sc.sweep[1] : none;
  auto stage = [&](const int4& h, int buf) {     // asynchronous copy of one sweep's rotations into ccs[buf]
    for (int t = tid; t < h.z; t += nt) cp_async16(&ccs[buf][t], sc.cs + h.x + t);
    asm volatile("cp.async.commit_group;\n" ::);
  };
  stage(h_cur, 0);
  if (tid < nr) {                                          // slab in (rows of Q from q_rows_kernel), thread = row
    const double* src = Z + (size_t)(row0 + tid);
    double* dst = slab + (size_t)tid * ld;
#pragma unroll 8
    for (int c = 0; c < n; ++c) dst[c] = src[(size_t)n * c];
  }
  // rotation t of a sweep acts on the column pair (m-1-t, m-t)
  for (int sw = 0; sw < nsweep; ++sw) {
    const int m = h_cur.y, cnt = h_cur.z;
    const double2* rc = ccs[sw & 1];
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();                      // sweep sw staged and visible; everyone is done with buffer (sw + 1) & 1
    const int4 h_nn = sw + 2 < nsweep ? sc.sweep[sw + 2] : none;
    stage(h_nxt, (sw + 1) & 1);
    if (live && cnt > 0) {
      // running column in a register (one DFMA on the dependent chain per rotation); the row entries of the next four
      // rotations are loaded before the current four are computed, so the shared-memory latency stays off the chain
      double carry = mine[m];
      int i = m - 1, t = 0;
      constexpr int U = 8;                 // rotations per software-pipelined group
      // group g + 1's operands -- row entries AND (c, s) pairs -- are loaded while group g is computed, so no
      // shared-memory latency sits between two groups; out-of-range prefetches are clamped to valid addresses
      double z[U], y[U];
      double2 r[U], rn[U];
      const int ngroups = cnt / U;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        z[u] = mine[max(i - u, 0)];
        r[u] = rc[min(u, kSeMaxN - 1)];
      }
      for (int g = 0; g < ngroups; ++g, t += U, i -= U) {
        const int tn = (g + 1 < ngroups) ? t + U : t;          // last group: harmless re-load of its own operands
        const int in = (g + 1 < ngroups) ? i - U : i;
#pragma unroll
        for (int u = 0; u < U; ++u) { y[u] = mine[in - u]; rn[u] = rc[tn + u]; }
        double a[U], b[U], o[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { a[u] = r[u].x * z[u]; b[u] = r[u].y * z[u]; }   // off the chain
#pragma unroll
        for (int u = 0; u < U; ++u) {
          o[u] = fma(r[u].x, carry, b[u]);
          carry = fma(-r[u].y, carry, a[u]);      // the dependent chain: one DFMA per rotation
        }
#pragma unroll
        for (int u = 0; u < U; ++u) mine[i + 1 - u] = o[u];
#pragma unroll
        for (int u = 0; u < U; ++u) { z[u] = y[u]; r[u] = rn[u]; }
      }
      for (; t < cnt; ++t, --i) {
        const double2 r = rc[t];
        const double z = mine[i];
        mine[i + 1] = fma(r.x, carry, r.y * z);
        carry = fma(-r.y, carry, r.x * z);
      }
      mine[i + 1] = carry;
    }
    h_cur = h_nxt;
    h_nxt = h_nn;
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  slab_store(slab, ld, Z, n, row0, nr);
}

// ascending order of the eigenvalues (ties by index) into perm[]; smem ints perm[n]
__device__ __forceinline__ void eig_sort_perm(const double* __restrict__ d, int n, int* perm) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = tid; k < n; k += nt) perm[k] = k;           // stays a valid index table even if an eigenvalue is NaN
  __syncthreads();
  for (int k = tid; k < n; k += nt) {
    const double lam = d[k];
    int below = 0;
    for (int t = 0; t < n; ++t) below += (d[t] < lam || (d[t] == lam && t < k)) ? 1 : 0;
    perm[below] = k;
  }
  __syncthreads();
}

// linearized_jacobians / residuals from (eigenvalues in the scratch, eigenvectors = columns of Z); smem: n doubles + n ints
__global__ void __launch_bounds__(256)
eig_prior_kernel(int n, const double* __restrict__ b_red, const double* __restrict__ Zbuf, const char* __restrict__ scratch,
                 double* __restrict__ LJ_all, double* __restrict__ LR_all, int32_t* __restrict__ rank_out,
                 int32_t* __restrict__ status, double eps) {
  extern __shared__ double smem[];
  const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  const double* br = b_red + (size_t)prob * n;
  const double* Z = Zbuf + (size_t)prob * n * n;
  double* LJ = LJ_all + (size_t)prob * n * n;
  double* LR = LR_all + (size_t)prob * n;
  const SymEigScratch sc = sym_eig_scratch(const_cast<char*>(scratch), n, prob);
  double* gb = smem;                                       // v_k . b_red
  int* perm = reinterpret_cast<int*>(smem + n);
  __shared__ int s_kept;
  if (tid == 0) s_kept = 0;
  for (int k = warp; k < n; k += nw) {
    const double* z = Z + (size_t)n * k;
    double a = 0.0;
    for (int i = lane; i < n; i += 32) a = fma(z[i], br[i], a);
    a = warp_sum(a);
    if (lane == 0) gb[k] = a;
  }
  eig_sort_perm(sc.d, n, perm);
  int kept_local = 0;
  for (int k = tid; k < n; k += nt) kept_local += (sc.d[k] > eps) ? 1 : 0;
  if (kept_local) atomicAdd(&s_kept, kept_local);
  // row r of linearized_jacobians (column-major n x n) = sqrt(lam) v^T for the r-th smallest eigenvalue, 0 if <= eps
  for (int r = warp; r < n; r += nw) {
    const int k = perm[r];
    const double lam = sc.d[k];
    const bool keep = lam > eps;
    const double sq = keep ? sqrt(lam) : 0.0;
    const double* z = Z + (size_t)n * k;
    for (int j = lane; j < n; j += 32) LJ[r + (size_t)n * j] = keep ? sq * z[j] : 0.0;
    if (lane == 0) LR[r] = keep ? gb[k] / sq : 0.0;
  }
  __syncthreads();
  if (tid == 0) {
    rank_out[prob] = s_kept;
    if (sc.meta[1] && status) atomicOr(status + prob, ISV_W_EIG_NOCONV);
  }
}

// unit-test hook: eigenvalues ascending + eigenvectors (column r of V = eigenvector of lam[r])
__global__ void __launch_bounds__(256)
eig_test_out_kernel(int n, const double* __restrict__ Zbuf, const char* __restrict__ scratch, double* __restrict__ lam_out,
                    double* __restrict__ V_out, int32_t* __restrict__ info) {
  extern __shared__ double smem[];
  const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const double* Z = Zbuf + (size_t)prob * n * n;
  const SymEigScratch sc = sym_eig_scratch(const_cast<char*>(scratch), n, prob);
  int* perm = reinterpret_cast<int*>(smem);
  eig_sort_perm(sc.d, n, perm);
  for (int r = warp; r < n; r += nw) {
    const int k = perm[r];
    const double* z = Z + (size_t)n * k;
    double* o = V_out + (size_t)prob * n * n + (size_t)n * r;
    for (int j = lane; j < n; j += 32) o[j] = z[j];
    if (lane == 0) lam_out[(size_t)prob * n + r] = sc.d[k];
  }
  if (tid == 0) { info[2 * prob] = sc.meta[1]; info[2 * prob + 1] = sc.meta[0]; }
}

}  // namespace isv
