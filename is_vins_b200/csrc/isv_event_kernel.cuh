// The latency path: ONE launch per MARGIN_OLD event (Estimator::MargForward(); Estimator::MargBackward(); called back to back
// from backendOptimization(), /root/reference/src/estimator.cpp:1541-1562).
//
// The batch path (isv_window_kernels.cuh) gives every window one warp per stage and runs the stages as five launches on three
// streams: right for thousands of windows, but a single estimator produces ONE event per keyframe, and then the time is the
// sum of the launch latencies plus the dependent chains  factor Jacobians -> backward  and  factor Jacobians / landmarks ->
// tail, each walked by one thread or one warp.  Here one CTA of eight warps owns the window and the chains run side by side:
//
//   warp 0      covariance Cholesky (needs only the pre-integration record) .. wait (named barrier 1) .. rest of MargBackward
//   warp 1      IMU factor Jacobian, the nine 3x3 sub-block positions dealt over nine lanes      -> arrive(1), landmark share
//   warp 2, 3   Jacobians of the new RelativePoseFactor / RollPitchFactor (one lane each)        -> arrive(1), landmark share
//   warp 4 .. 7 the four forward factor Jacobians (one lane each)                                -> landmark share
//   warp 1 .. 7 split the window's landmarks (forward_accum_body on a contiguous share each), park the partial Gram
//               triangles in shared memory, arrive(2); warp 7 waits on barrier 2, adds the partials in fixed order and runs
//               the forward tail.
//
// The factor-Jacobian record and the Gram triangles never leave shared memory.  Every device function called here is the one
// the batch kernels call, so the arithmetic is the same; only the summation order of the landmark Gram differs (7 partial
// sums instead of one), i.e. results agree with the batch path to rounding, not bit for bit.
//
// Inputs and outputs may live in MAPPED PINNED HOST memory (isv_marg_event): the kernel then reads the event straight over
// PCIe and writes the recovered factors straight back, and publishes `done_seq` to `done_flag` (system-scope fence first) so
// that the host can spin on a word of its own memory instead of paying two copies and a stream synchronisation.  For that
// reason status bits are collected in shared memory (no atomics on host memory) and stored once at the end.
#pragma once
#include "isv_window_kernels.cuh"

namespace isv {

constexpr int kEvWarps = 8;
constexpr int kEvThreads = 32 * kEvWarps;
constexpr int kEvAccWarps = 7;   // warps 1 .. 7
// shared-memory map (doubles; every offset even = 16-byte aligned)
constexpr int kEvFJ = 0;                                            // factor-Jacobian record
constexpr int kEvBwd = kEvFJ + kFJ;                                 // work area of the backward warp
constexpr int kEvTail = kEvBwd + kBwdSmemPerWarp;                   // work area of the tail
constexpr int kEvAcc = kEvTail + kFwdSmemPerWarp;                   // constants + reduction staging of the landmark warps
constexpr int kEvPart = kEvAcc + kEvAccWarps * kAccSmemPerWarp;     // partial Gram triangles [7][42]
constexpr int kEvGram = kEvPart + kEvAccWarps * 42;                 // their sum [42]
constexpr int kEvStat = kEvGram + 42;                               // status word
constexpr int kEvSmemDoubles = kEvStat + 2;
static_assert(kEvSmemDoubles % 2 == 0, "the staged event block starts 16-byte aligned");
constexpr int kEvStageMaxDoubles = (200 * 1024) / 8 - kEvSmemDoubles;   // what fits next to the work areas in one SM
static_assert((kFJ % 2 == 0) && (kBwdSmemPerWarp % 2 == 0) && (kFwdSmemPerWarp % 2 == 0) && (kAccSmemPerWarp % 2 == 0),
              "16-byte alignment of the shared-memory map");

// ---- bulk-async staging of one event's inputs (isv_marg_event) --------------------------------------------------------
// The event block is one contiguous, 16-byte aligned region [lm_offset | records | landmark components]: one thread hands it
// to the TMA engine as ONE cp.async.bulk (SASS UBLKCP) that completes on an mbarrier, and every chain then starts from
// shared memory instead of paying its own first-touch round trip to HBM -- or, on the zero-copy route, over PCIe to the
// host's pinned block (~2 us each, several per chain).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

// stage_doubles > 0: the first stage_doubles doubles at in.lm_offset (the packed event block of isv_marg_event) are staged
// into shared memory behind the kernel's work areas and every input pointer inside that range is re-based onto the copy.
// lam_comp: position of the inverse-depth component in lm_obs (5 in the ABI's layout; 2 / 3 in the packed event block).
template <bool ZONE, bool ISO>
__global__ void __launch_bounds__(kEvThreads, 1)
marg_event_fused_kernel(isv_batch_in in_arg, isv_batch_out out, DevCfg cfg, int32_t* done_flag, int32_t done_seq,
                        long long* stamps, int stage_doubles, int lam_comp) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int win = blockIdx.x;
  isv_batch_in in = in_arg;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + kEvStat + 1);
  if (stage_doubles > 0 && threadIdx.x == 0) {
    mbar_init(mbar, 1);
    bulk_g2s(smem + kEvSmemDoubles, in_arg.lm_offset, (uint32_t)stage_doubles * 8u, mbar);
  }
  // profiling aid (isv_test_fused_stamps): SM cycle counter at the phase boundaries of every warp of window 0
  int n_stamp = 0;
  auto stamp = [&]() {
    if (stamps && win == 0 && lane == 0) stamps[8 * warp + n_stamp] = clock64();
    ++n_stamp;
  };
  stamp();
  double* FJ = smem + kEvFJ;
  int32_t* sstat = reinterpret_cast<int32_t*>(smem + kEvStat);
  // the IMU Jacobian record is sparse: zero-fill it, imu_jacobians writes the non-zero blocks
  for (int i = threadIdx.x; i < 450; i += kEvThreads) FJ[kFJ_IMU + i] = 0.0;
  if (threadIdx.x == 0) *sstat = 0;
  __syncthreads();
  if (stage_doubles > 0) {
    mbar_wait(mbar, 0);
    const char* base = reinterpret_cast<const char*>(in_arg.lm_offset);
    const char* copy = reinterpret_cast<const char*>(smem + kEvSmemDoubles);
    const long long span = (long long)stage_doubles * 8;
    auto rb = [&](const double* p) -> const double* {
      const long long o = reinterpret_cast<const char*>(p) - base;
      return (p && o >= 0 && o < span) ? reinterpret_cast<const double*>(copy + o) : p;
    };
    in.lm_offset = reinterpret_cast<const int64_t*>(copy);
    in.lm_obs = rb(in.lm_obs);
    in.pose_fwd = rb(in.pose_fwd); in.ex_pose = rb(in.ex_pose); in.prior_se3 = rb(in.prior_se3);
    in.prior_rel = rb(in.prior_rel); in.prior_rp = rb(in.prior_rp);
    in.pose_bwd = rb(in.pose_bwd); in.sb_bwd = rb(in.sb_bwd); in.prior_vb = rb(in.prior_vb); in.preint = rb(in.preint);
  }
  stamp();

  if (warp == 0) {
    backward_body<128>(in, out, cfg, win, lane, smem + kEvBwd, FJ, sstat, nullptr, (stamps && win == 0) ? stamps + 64 : nullptr);
    stamp();
  } else {
    if (warp == 1) {
      const double* pose_i = in.pose_bwd + (size_t)win * 14;
      const double* pose_j = pose_i + 7;
      const double* sb_i = in.sb_bwd + (size_t)win * 18;
      const double* sb_j = sb_i + 9;
      const double* pre = in.preint + (size_t)win * ISV_PREINT_REC;
      if (lane < 9) imu_jacobians(pose_i, sb_i, pose_j, sb_j, pre, cfg.g, FJ + kFJ_IMU, 1, 15, 21, 0, 6, nullptr, lane, 9, 30);
      else if (lane == 9 && (nonunit(pose_i) || nonunit(pose_j))) atomicOr(sstat, ISV_W_NONUNIT_QUAT);
    } else if (lane == 0) {
      // warp 2 -> task 5, warp 3 -> task 6 (backward), warps 4 .. 7 -> tasks 0 .. 3 (forward)
      factor_jac_task(in, out, FJ, cfg, win, warp < 4 ? warp + 3 : warp - 4, sstat);
    }
    __syncwarp();
    stamp();
    if (warp < 4) {
      __threadfence_block();
      named_bar_arrive(kFusedBarBwd, 128);
    }
    // landmark share of this warp
    const long long lm0 = in.lm_offset[win];
    const int L = (int)(in.lm_offset[win + 1] - lm0);
    const int a = warp - 1;
    const int per = (L + kEvAccWarps - 1) / kEvAccWarps;
    const int b0 = min(L, a * per);
    const int cnt = min(per, L - b0);
    forward_accum_body<ZONE, ISO>(in, cfg, win, lane, smem + kEvAcc + a * kAccSmemPerWarp, lm0 + b0, cnt, smem + kEvPart + 42 * a,
                                  sstat, lam_comp);
    stamp();
    if (warp < kEvWarps - 1) {
      __threadfence_block();
      named_bar_arrive(kFusedBarFwd, 32 * kEvAccWarps);
    } else {
      named_bar_sync(kFusedBarFwd, 32 * kEvAccWarps);
      stamp();
      double* G = smem + kEvGram;
      for (int t = lane; t < 42; t += 32) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < kEvAccWarps; ++w) acc += smem[kEvPart + 42 * w + t];   // fixed order: deterministic
        G[t] = acc;
      }
      __syncwarp();
      forward_tail_body(in, out, cfg, win, lane, smem + kEvTail, G, FJ, sstat, nullptr);
      stamp();
    }
  }
  // every store of the recovered factors is made visible system-wide (the outputs may be mapped host memory) before the
  // status word -- and, for a single event, the completion flag -- is published
  __threadfence_system();
  __syncthreads();
  if (stamps && win == 0 && lane == 0) stamps[8 * warp + 7] = clock64();
  if (threadIdx.x == 0) {
    if (out.status) out.status[win] = *sstat;
    if (done_flag) {
      __threadfence_system();
      *reinterpret_cast<volatile int32_t*>(done_flag) = done_seq;
    }
  }
}

}  // namespace isv
