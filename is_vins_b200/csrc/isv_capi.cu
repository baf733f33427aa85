// C ABI of the B200-native IS-VINS marginalization backend (include/isv_capi.h).
// Host side only: argument checks, device scratch owned by the handle, H2D/D2H staging for the
// host-pointer entry points, kernel launches.  No CPU fallback anywhere: without a usable sm_100
// device every compute entry point returns ISV_ERR_CUDA.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <vector>

#include "isv_eval_kernels.cuh"
#include "isv_init_kernel.cuh"
#include "isv_marg_generic.cuh"
#include "isv_sym_eig.cuh"
#include "isv_preint_kernel.cuh"
#include "isv_seq_kernels.cuh"
#include "isv_window_kernels.cuh"
#include "isv_event_kernel.cuh"
#include "isv_forensic.cuh"

using namespace isv;

struct isv_handle {
  int device;
  cudaStream_t own_stream;
  cudaStream_t stream;
  cudaStream_t copy_stream;
  cudaStream_t aux[4];   // fork/join side streams: [0..1] isv_eval_problem and the backward chain (per slot), [2..3] forward factor Jacobians
  cudaEvent_t aux_ev[3];
  isv_config cfg;
  DevCfg dcfg;
  int64_t launches;
  // device mirror of a host batch (grow-only)
  char* dbuf;
  size_t dbuf_bytes;
  char* pinned;
  size_t pinned_bytes;
  char* mapped;          // isv_marg_event: mapped pinned block the fused kernel reads / writes over PCIe (zero-copy)
  char* mapped_dev;      //   its device alias
  size_t mapped_bytes;
  int32_t event_seq;     //   completion-flag sequence number
  int fused_max;         // ISV_TUNE_FUSED_MAX_WINDOWS
  int acc_persist;       // ISV_TUNE_ACC_PERSIST: persistent landmark-kernel warps per SM (0 = one CTA per window)
  int n_sm;
  int32_t* counters;     // device: work counters of the persistent kernels, one 16-int slot per launching stream
  int event_mode;        // ISV_TUNE_EVENT_MODE
  char* eig;             // eigensolver scratch of the generic engine (tridiagonal + rotation log per problem), grow-only
  size_t eig_bytes;
  double* gram;          // [n][42 + kFJ] scratch handed between the kernels of one batch, grow-only:
  size_t gram_bytes;     //   landmark Gram triangles (forward stage 1 -> 2) and the factor-Jacobian records
  cudaEvent_t jac_ev[12]; // fork / join events of launch_batch, three per launching slot
  cudaStream_t pipe[2];   // third and fourth pipeline stream of the chunked host path (slots 2, 3)
  cudaStream_t h2d_stream;   // every H2D copy of the chunked host path, in chunk order
  cudaStream_t bwd_stream;   // the backward half of the host path (it needs no landmarks: once, for the whole batch)
  cudaEvent_t chunk_ev[16];  // "chunk c is on the device"
  cudaStream_t fork[8];   // side streams of launch_batch: [2 slot] backward chain, [2 slot + 1] forward factor Jacobians
  // isv_marg_window_batch: the ~15 runtime calls of one launch_batch (fork / join over three streams) replayed as one
  // CUDA graph when the same buffers come back (a server loop re-fills the same device batch every step)
  struct BatchGraph { isv_batch_in in; isv_batch_out out; int which; cudaStream_t stream; double* gram; cudaGraphExec_t exec; int launches; };
  BatchGraph bgraph[4];
  int bgraph_next, bgraph_miss;   // after 8 consecutive misses (a caller that never repeats a batch) capturing stops
  cudaEvent_t ev[4];
};

constexpr int kFusedMaxWindows = 148;   // one CTA per SM: see launch_fused
constexpr int kAccPersistPerSm = 0;     // persistent landmark-kernel warps per SM (see marg_forward_accum_kernel): OFF by default,
constexpr int kAccPersistMinWindows = 2368;   // measured without gain; when on, used for batches of at least this many windows
// ABI 4 record layouts (include/isv_capi.h ISV_IN_TRI_RECORDS / ISV_OUT_TRI_RECORDS)
static const TriLayout kTriSe3 = {2, {0, 1, 0, 0}, {12, 6, 0, 0}};       // 48 <-> 33 (also the RelativePoseFactor record)
static const TriLayout kTriVb = {2, {0, 1, 0, 0}, {9, 9, 0, 0}};         // 90 <-> 54
static const TriLayout kTriRpIn = {2, {0, 1, 0, 0}, {1, 2, 0, 0}};       // 5 <-> 4
static const TriLayout kTriPg = {4, {0, 1, 1, 0}, {12, 6, 6, 5}};        // 89 <-> 59
static const TriLayout kTriRpOut = {2, {0, 1, 0, 0}, {9, 2, 0, 0}};      // 13 <-> 12
template <bool PACK>
static void tri_launch(isv_handle* h, cudaStream_t s, size_t n, const double* src, double* dst, const TriLayout& lay) {
  if (n == 0) return;
  const long long total = (long long)n * tri_full_len(lay);
  const int grid = (int)(total / 256 + 1 < 1184 ? total / 256 + 1 : 1184);
  tri_records_kernel<PACK><<<grid, 256, 0, s>>>((long long)n, src, dst, lay);
  ++h->launches;
}

// the landmark kernel's eight instantiations: pts_i.z == 1 promised / sqrt_info = c I / pts_i.xy as FP32
static const void* accum_kernel_variant(bool zone, bool iso, bool xyf) {
  static const void* const k[8] = {
      (const void*)marg_forward_accum_kernel<false, false, false>, (const void*)marg_forward_accum_kernel<true, false, false>,
      (const void*)marg_forward_accum_kernel<false, true, false>,  (const void*)marg_forward_accum_kernel<true, true, false>,
      (const void*)marg_forward_accum_kernel<false, false, true>,  (const void*)marg_forward_accum_kernel<true, false, true>,
      (const void*)marg_forward_accum_kernel<false, true, true>,   (const void*)marg_forward_accum_kernel<true, true, true>};
  return k[(zone ? 1 : 0) | (iso ? 2 : 0) | (xyf ? 4 : 0)];
}


#define ISV_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      fprintf(stderr, "[isv_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return ISV_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

extern "C" {

void isv_default_config(isv_config* c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->alpha = 0.1;                 // config/euroc_config.yaml:86
  c->proj_sqrt_info[0] = 460.0;   // yaml:83, src/estimator.cpp:35
  c->proj_sqrt_info[3] = 460.0;
  c->g[2] = 9.81007;              // yaml g_norm, src/parameters.cpp:96
  c->acc_n = 0.22627;             // yaml:57-60
  c->gyr_n = 0.003988;
  c->acc_w = 0.001;
  c->gyr_w = 0.0001;
  c->vo_size = 8;                 // include/parameters.h:35
  c->all_buf_size = 18;           // include/parameters.h:40
  c->qr_rank_eps_log10 = -16;     // src/estimator.cpp:8
}

int isv_abi_version(void) { return ISV_ABI_VERSION; }

const char* isv_status_string(isv_status s) {
  switch (s) {
    case ISV_OK: return "ISV_OK";
    case ISV_ERR_BAD_ARG: return "ISV_ERR_BAD_ARG";
    case ISV_ERR_CUDA: return "ISV_ERR_CUDA";
    case ISV_ERR_ALLOC: return "ISV_ERR_ALLOC";
  }
  return "ISV_ERR_UNKNOWN";
}

isv_status isv_create(const isv_config* cfg, int device, isv_handle** out) {
  if (!cfg || !out) return ISV_ERR_BAD_ARG;
  if (cfg->vo_size < 2 || !(cfg->alpha >= 0.0)) return ISV_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  ISV_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return ISV_ERR_CUDA;
  cudaDeviceProp prop;
  ISV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    fprintf(stderr, "[isv_b200] device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major,
            prop.minor);
    return ISV_ERR_CUDA;
  }
  ISV_CUDA(cudaSetDevice(device));
  isv_handle* h = new (std::nothrow) isv_handle();
  if (!h) return ISV_ERR_ALLOC;
  memset(h, 0, sizeof(*h));
  h->device = device;
  h->cfg = *cfg;
  h->dcfg.alpha = cfg->alpha;
  for (int i = 0; i < 4; ++i) h->dcfg.ps[i] = cfg->proj_sqrt_info[i];
  for (int i = 0; i < 3; ++i) h->dcfg.g[i] = cfg->g[i];
  h->dcfg.qr_threshold = pow(10.0, (double)cfg->qr_rank_eps_log10);
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return ISV_ERR_CUDA;
  }
  for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming);
  for (int i = 0; i < 4; ++i) cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking);
  for (int i = 0; i < 3; ++i) cudaEventCreateWithFlags(&h->aux_ev[i], cudaEventDisableTiming);
  for (int i = 0; i < 12; ++i) cudaEventCreateWithFlags(&h->jac_ev[i], cudaEventDisableTiming);
  for (int i = 0; i < 2; ++i) cudaStreamCreateWithFlags(&h->pipe[i], cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->bwd_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 16; ++i) cudaEventCreateWithFlags(&h->chunk_ev[i], cudaEventDisableTiming);
  for (int i = 0; i < 8; ++i) cudaStreamCreateWithFlags(&h->fork[i], cudaStreamNonBlocking);
  h->stream = h->own_stream;
  {
    const char* e = getenv("ISV_FUSED_MAX");
    h->fused_max = e ? atoi(e) : kFusedMaxWindows;
    e = getenv("ISV_ACC_PERSIST");
    h->acc_persist = e ? atoi(e) : kAccPersistPerSm;
    h->n_sm = prop.multiProcessorCount;
    if (cudaMalloc(&h->counters, 64 * sizeof(int32_t)) != cudaSuccess) { cudaGetLastError(); h->counters = nullptr; h->acc_persist = 0; }
    e = getenv("ISV_EVENT_MODE");
    h->event_mode = e ? atoi(e) : 0;
    if (h->event_mode < 0 || h->event_mode > 2) h->event_mode = 0;
  }
  {
    const int sm = (int)(kAccWarps * kAccSmemPerWarp * sizeof(double));
    for (int v = 0; v < 8; ++v) cudaFuncSetAttribute(accum_kernel_variant(v & 1, v & 2, v & 4), cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
  }
  cudaFuncSetAttribute(marg_forward_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(kWarpsPerCta * kFwdSmemPerWarp * sizeof(double)));
  cudaFuncSetAttribute(marg_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(kWarpsPerCta * kBwdSmemPerWarp * sizeof(double)));
  cudaFuncSetAttribute(preintegrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(kWarpsPerCta * kPreSmemPerWarp * sizeof(double)));
  cudaFuncSetAttribute(marg_event_fused_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)((kEvSmemDoubles + kEvStageMaxDoubles) * sizeof(double)));
  cudaFuncSetAttribute(marg_event_fused_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)((kEvSmemDoubles + kEvStageMaxDoubles) * sizeof(double)));
  *out = h;
  return ISV_OK;
}

void isv_destroy(isv_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->dbuf) cudaFree(h->dbuf);
  if (h->gram) cudaFree(h->gram);
  if (h->eig) cudaFree(h->eig);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->mapped) cudaFreeHost(h->mapped);
  if (h->counters) cudaFree(h->counters);
  for (int i = 0; i < 4; ++i)
    if (h->bgraph[i].exec) cudaGraphExecDestroy(h->bgraph[i].exec);
  for (int i = 0; i < 4; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < 4; ++i)
    if (h->aux[i]) cudaStreamDestroy(h->aux[i]);
  for (int i = 0; i < 3; ++i)
    if (h->aux_ev[i]) cudaEventDestroy(h->aux_ev[i]);
  for (int i = 0; i < 12; ++i)
    if (h->jac_ev[i]) cudaEventDestroy(h->jac_ev[i]);
  for (int i = 0; i < 2; ++i)
    if (h->pipe[i]) cudaStreamDestroy(h->pipe[i]);
  if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
  if (h->bwd_stream) cudaStreamDestroy(h->bwd_stream);
  for (int i = 0; i < 16; ++i)
    if (h->chunk_ev[i]) cudaEventDestroy(h->chunk_ev[i]);
  for (int i = 0; i < 8; ++i)
    if (h->fork[i]) cudaStreamDestroy(h->fork[i]);
  cudaStreamDestroy(h->own_stream);
  cudaStreamDestroy(h->copy_stream);
  delete h;
}

void* isv_stream(isv_handle* h) { return h ? (void*)h->stream : nullptr; }

isv_status isv_set_stream(isv_handle* h, void* s) {
  if (!h) return ISV_ERR_BAD_ARG;
  h->stream = s ? (cudaStream_t)s : h->own_stream;
  return ISV_OK;
}

isv_status isv_synchronize(isv_handle* h) {
  if (!h) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  ISV_CUDA(cudaStreamSynchronize(h->stream));
  return ISV_OK;
}

int64_t isv_launch_count(const isv_handle* h) { return h ? h->launches : 0; }

isv_status isv_set_tuning(isv_handle* h, int knob, int value) {
  if (!h) return ISV_ERR_BAD_ARG;
  if (knob == ISV_TUNE_FUSED_MAX_WINDOWS && value >= 0) { h->fused_max = value; return ISV_OK; }
  if (knob == ISV_TUNE_EVENT_MODE && value >= 0 && value <= 2) { h->event_mode = value; return ISV_OK; }
  if (knob == ISV_TUNE_ACC_PERSIST && value >= 0 && value <= 16) { h->acc_persist = h->counters ? value : 0; return ISV_OK; }
  return ISV_ERR_BAD_ARG;
}

// ---- index maps -------------------------------------------------------------------------------
int isv_order_map_init(int V, int32_t* out) {
  if (V < 2 || !out) return 0;
  int idx = 0, b = 0;
  for (int i = 0; i < V; ++i) { out[2 * b] = idx; out[2 * b + 1] = 6; idx += 6; ++b; }
  out[2 * b] = idx; out[2 * b + 1] = 9; idx += 9; ++b;
  for (int i = 0; i < V - 1; ++i) { out[2 * b] = idx; out[2 * b + 1] = 9; idx += 9; ++b; }
  return b;
}

int isv_order_map_forward(int L, int32_t* out) {
  if (L < 0 || !out) return 0;
  out[0] = 0; out[1] = 6; out[2] = 6; out[3] = 6;
  for (int k = 0; k < L; ++k) { out[4 + 2 * k] = 12 + k; out[5 + 2 * k] = 1; }
  return 2 + L;
}

int isv_order_map_backward(int V, int32_t* out) {
  (void)V;
  if (!out) return 0;
  const int32_t m[8] = {0, 6, 6, 9, 15, 6, 21, 9};
  memcpy(out, m, sizeof(m));
  return 4;
}

// ---- batched device entry point ---------------------------------------------------------------
static isv_status check_batch(const isv_batch_in* in, const isv_batch_out* out, int which, bool allow_tri = false) {
  constexpr int kAll = ISV_RUN_BOTH | ISV_RUN_FORWARD_STAGE1 | ISV_RUN_FORWARD_STAGE2 | ISV_RUN_FACTOR_JAC | ISV_RUN_BACKWARD_STAGE2;
  if (!in || !out || in->n_windows < 0 || (which & ~kAll) || which == 0) return ISV_ERR_BAD_ARG;
  if (!out->rank) return ISV_ERR_BAD_ARG;
  if (which & (ISV_RUN_FORWARD | ISV_RUN_FORWARD_STAGE1 | ISV_RUN_FORWARD_STAGE2 | ISV_RUN_FACTOR_JAC)) {
    if (!in->lm_offset || !in->pose_fwd || !in->ex_pose || !in->prior_se3 || !in->prior_rel || !out->se3_out ||
        !out->pg_out)
      return ISV_ERR_BAD_ARG;
    if (!in->lm_obs && in->lm_stride != 0) return ISV_ERR_BAD_ARG;
  }
  if (in->flags & ~(ISV_IN_PTS_I_Z_ONE | (allow_tri ? (ISV_IN_TRI_RECORDS | ISV_OUT_TRI_RECORDS) : 0))) return ISV_ERR_BAD_ARG;
  if (which & (ISV_RUN_BACKWARD | ISV_RUN_BACKWARD_STAGE2 | ISV_RUN_FACTOR_JAC)) {
    if (!in->pose_bwd || !in->sb_bwd || !in->prior_vb || !out->rel_out || !out->vb_out || !out->rp_out)
      return ISV_ERR_BAD_ARG;
    // the pre-integration record itself, or the raw samples it is rebuilt from (ABI 2)
    if (!in->preint && !(in->imu_init && in->imu_k_max >= 0 && (in->imu_raw || in->imu_k_max == 0))) return ISV_ERR_BAD_ARG;
  }
  return ISV_OK;
}

// scratch: device memory [n_windows][kScratchPerWindow]: the landmark Gram triangles (42) handed from the
// landmark kernel to the tail kernel, then the factor-Jacobian records (kFJ) of marg_factor_jac_kernel
// and (when the caller hands raw IMU samples instead of the record) the pre-integration records
constexpr size_t kScratchPerWindow = 42 + kFJ + ISV_PREINT_REC;

struct DbgStores { double* lamda_prior_fwd; double* g_bwd; };

// occupancy experiments (measurement only): extra dynamic shared memory per CTA of each window kernel
static size_t exp_smem(const char* name) {
  const char* e = getenv(name);
  return e ? (size_t)atol(e) : 0;
}

static isv_status launch_batch(isv_handle* h, const isv_batch_in* in_arg, const isv_batch_out* out, int which,
                               cudaStream_t stream, double* scratch, DbgStores dbg = DbgStores{nullptr, nullptr},
                               bool zero_status = true) {
  static const size_t x_acc = exp_smem("ISV_EXP_ACC_SMEM"), x_tail = exp_smem("ISV_EXP_TAIL_SMEM"), x_bwd = exp_smem("ISV_EXP_BWD_SMEM");
  static const bool x_once = [] {
    if (x_acc) {
      const int sm = (int)(kAccWarps * kAccSmemPerWarp * sizeof(double) + x_acc);
      for (int v = 0; v < 8; ++v) cudaFuncSetAttribute(accum_kernel_variant(v & 1, v & 2, v & 4), cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    }
    if (x_tail) cudaFuncSetAttribute(marg_forward_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(kWarpsPerCta * kFwdSmemPerWarp * sizeof(double) + x_tail));
    if (x_bwd) cudaFuncSetAttribute(marg_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kWarpsPerCta * kBwdSmemPerWarp * sizeof(double) + x_bwd));
    return true;
  }();
  (void)x_once;
  const int n = in_arg->n_windows;
  if (n == 0) return ISV_OK;
  if (!scratch) return ISV_ERR_BAD_ARG;
  double* gram = scratch;
  double* fj = scratch + (size_t)n * 42;
  // raw IMU in, no record: preintegrate_kernel fills the scratch record first, on the backward chain's stream
  isv_batch_in in_local = *in_arg;
  const bool need_preint = !in_arg->preint && (which & (ISV_RUN_BACKWARD | ISV_RUN_FACTOR_JAC | ISV_RUN_BACKWARD_STAGE2));
  if (need_preint) in_local.preint = scratch + (size_t)n * (42 + kFJ);
  const isv_batch_in* in = &in_local;
  const bool stage1 = which & (ISV_RUN_FORWARD | ISV_RUN_FORWARD_STAGE1);
  const bool stage2 = which & (ISV_RUN_FORWARD | ISV_RUN_FORWARD_STAGE2);
  const bool jac_fwd = which & (ISV_RUN_FORWARD | ISV_RUN_FACTOR_JAC);
  const bool jac_bwd = which & (ISV_RUN_BACKWARD | ISV_RUN_FACTOR_JAC);
  const bool bwd = which & (ISV_RUN_BACKWARD | ISV_RUN_BACKWARD_STAGE2);
  const bool stage1_ = which & (ISV_RUN_FORWARD | ISV_RUN_FORWARD_STAGE1);
  const bool persist = stage1_ && h->acc_persist > 0 && h->counters && n >= kAccPersistMinWindows;
  const int slot = (stream == h->copy_stream) ? 1 : (stream == h->pipe[0] ? 2 : (stream == h->pipe[1] ? 3 : 0));
  int32_t* counter = persist ? h->counters + 16 * slot : nullptr;
  if ((out->status && zero_status) || persist) {
    int32_t dummy_n = (out->status && zero_status) ? n : 0;
    zero_i32_kernel<<<(n + 255) / 256, 256, 0, stream>>>(out->status, dummy_n, counter, persist ? 1 : 0);
    ++h->launches;
  }
  const int grid = (n + kWarpsPerCta - 1) / kWarpsPerCta;
  // Three independent chains: [landmark kernel] on the caller's stream, [forward factor Jacobians] and
  // [backward factor Jacobians -> marg_backward_kernel] on side streams.  The factor-Jacobian launches are
  // pure latency (one thread per window and factor, ~1.5 % issue utilisation): started together with the
  // FP64-bound landmark kernel they hide behind it; the tail kernel joins the forward Jacobians, the end of
  // the call joins the backward chain.
  const bool fork_b = (which & ISV_RUN_BACKWARD) && (which & ISV_RUN_FORWARD);
  const bool fork_f = jac_fwd && stage1;
  cudaEvent_t* ev = h->jac_ev + 3 * slot;   // [0] fork point, [1] backward chain done, [2] forward Jacobians done
  cudaStream_t bs = fork_b ? h->fork[2 * slot] : stream;
  cudaStream_t fs = fork_f ? h->fork[2 * slot + 1] : stream;
  if (fork_b || fork_f) ISV_CUDA(cudaEventRecord(ev[0], stream));
  if (fork_b) ISV_CUDA(cudaStreamWaitEvent(bs, ev[0], 0));
  if (fork_f) ISV_CUDA(cudaStreamWaitEvent(fs, ev[0], 0));
  if (jac_fwd) {
    marg_factor_jac_kernel<<<dim3((n + 127) / 128, 4), 128, 0, fs>>>(*in, *out, fj, h->dcfg, 0);
    ++h->launches;
    if (fork_f) ISV_CUDA(cudaEventRecord(ev[2], fs));
  }
  if (need_preint && jac_bwd) {
    NoiseCfg nz{h->cfg.acc_n, h->cfg.gyr_n, h->cfg.acc_w, h->cfg.gyr_w};
    preintegrate_kernel<<<grid, kThreads, kWarpsPerCta * kPreSmemPerWarp * sizeof(double), bs>>>(
        n, in->imu_k_max, in->imu_count, in->imu_raw, in->imu_init, const_cast<double*>(in->preint), nz);
    ++h->launches;
  }
  if (jac_bwd) {
    // the IMU Jacobian record is sparse: zero-fill it, the kernel writes the non-zero blocks
    {
      const long long tot = (long long)n * 450;
      const int zgrid = (int)(tot / 1024 < 1184 ? (tot + 1023) / 1024 : 1184);   // <= 8 CTAs per SM, 4 elements per thread and pass
      zero_rows_kernel<<<zgrid, 256, 0, bs>>>(fj + kFJ_IMU, n, 450, kFJ);
      ++h->launches;
    }
    marg_factor_jac_kernel<<<dim3((n + 127) / 128, 3), 128, 0, bs>>>(*in, *out, fj, h->dcfg, 4);
    ++h->launches;
  }
  if (stage1) {
    const int agrid = (n + kAccWarps - 1) / kAccWarps;
    const size_t asm_ = kAccWarps * kAccSmemPerWarp * sizeof(double) + x_acc;
    const bool zone = (in->flags & ISV_IN_PTS_I_Z_ONE) != 0;
    // ProjectionFactor::sqrt_info = c I (always, in the reference: src/estimator.cpp:35) takes the leaner chain
    static const bool no_iso = getenv("ISV_NO_ISO") != nullptr;   // A/B switch for measurements
    const bool iso = !no_iso && h->dcfg.ps[1] == 0.0 && h->dcfg.ps[2] == 0.0 && h->dcfg.ps[0] == h->dcfg.ps[3];
    const bool xyf = in->lm_xy_f32 != nullptr;
    const void* ak = accum_kernel_variant(zone, iso, xyf);
    isv_batch_in a_in = *in;
    int32_t* a_status = out->status;
    DevCfg a_cfg = h->dcfg;
    void* args[5] = {&a_in, &gram, &a_status, &a_cfg, &counter};
    const int pgrid = persist ? h->acc_persist * h->n_sm / kAccWarps : agrid;
    ISV_CUDA(cudaLaunchKernel(ak, dim3(pgrid < agrid ? pgrid : agrid), dim3(32 * kAccWarps), args, asm_, stream));
    ++h->launches;
  }
  if (bwd) {
    marg_backward_kernel<<<grid, kThreads, kWarpsPerCta * kBwdSmemPerWarp * sizeof(double) + x_bwd, bs>>>(*in, *out, fj, h->dcfg,
                                                                                                 h->cfg.vo_size, dbg.g_bwd);
    ++h->launches;
  }
  if (fork_b) ISV_CUDA(cudaEventRecord(ev[1], bs));
  if (fork_f) ISV_CUDA(cudaStreamWaitEvent(stream, ev[2], 0));
  if (stage2) {
    marg_forward_tail_kernel<<<grid, kThreads, kWarpsPerCta * kFwdSmemPerWarp * sizeof(double) + x_tail, stream>>>(
        *in, *out, gram, fj, h->dcfg, dbg.lamda_prior_fwd);
    ++h->launches;
  }
  if (fork_b) ISV_CUDA(cudaStreamWaitEvent(stream, ev[1], 0));
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

// ---- the latency path: one fused launch, one CTA per window (isv_event_kernel.cuh) --------------------------------------
// Eligible: both halves wanted, the pre-integration record given, ProjectionFactor::sqrt_info = c I (always, in the reference:
// src/estimator.cpp:35).  Measured on B200 (profiles/r02q_*): wins up to kFusedMaxWindows windows, where the grid stops
// fitting one CTA per SM; beyond that the warp-per-window batch kernels have the higher throughput.
static bool fused_eligible(const isv_handle* h, const isv_batch_in* in, int which) {
  const bool iso = h->dcfg.ps[1] == 0.0 && h->dcfg.ps[2] == 0.0 && h->dcfg.ps[0] == h->dcfg.ps[3];
  return which == ISV_RUN_BOTH && in->preint && !in->lm_xy_f32 && iso && in->n_windows >= 1 && in->n_windows <= h->fused_max;
}
static isv_status launch_fused(isv_handle* h, const isv_batch_in* in, const isv_batch_out* out, cudaStream_t stream,
                               int32_t* done_flag, int32_t done_seq, long long* stamps = nullptr, int stage_doubles = 0,
                               int lam_comp = 5) {
  const size_t sm = (size_t)(kEvSmemDoubles + stage_doubles) * sizeof(double);
  if (in->flags & ISV_IN_PTS_I_Z_ONE)
    marg_event_fused_kernel<true, true><<<in->n_windows, kEvThreads, sm, stream>>>(*in, *out, h->dcfg, done_flag, done_seq, stamps,
                                                                                   stage_doubles, lam_comp);
  else
    marg_event_fused_kernel<false, true><<<in->n_windows, kEvThreads, sm, stream>>>(*in, *out, h->dcfg, done_flag, done_seq, stamps,
                                                                                    stage_doubles, lam_comp);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

isv_status isv_marg_window_batch(isv_handle* h, const isv_batch_in* in, const isv_batch_out* out, int which) {
  if (!h) return ISV_ERR_BAD_ARG;
  isv_status st = check_batch(in, out, which);
  if (st != ISV_OK) return st;
  ISV_CUDA(cudaSetDevice(h->device));
  if (fused_eligible(h, in, which)) return launch_fused(h, in, out, h->stream, nullptr, 0);
  double* gram = nullptr;
  {
    const size_t need = (size_t)in->n_windows * kScratchPerWindow * sizeof(double);
    if (h->gram_bytes < need) {
      if (h->gram) {
        ISV_CUDA(cudaStreamSynchronize(h->stream));
        ISV_CUDA(cudaFree(h->gram));
        h->gram = nullptr;
        h->gram_bytes = 0;
      }
      if (cudaMalloc(&h->gram, need + need / 4 + 256) != cudaSuccess) {
        cudaGetLastError();
        return ISV_ERR_ALLOC;
      }
      h->gram_bytes = need + need / 4 + 256;
    }
    gram = h->gram;
  }
  // same buffers, same stages, same stream as a previous call: replay its graph (the GPU then schedules the forked chains
  // itself instead of waiting for the host to issue each of them: it matters when a batch is a fraction of a wave)
  cudaStream_t s = h->stream;
  for (int i = 0; i < 4; ++i) {
    isv_handle::BatchGraph& g = h->bgraph[i];
    if (g.exec && g.which == which && g.stream == s && g.gram == gram && memcmp(&g.in, in, sizeof(*in)) == 0 &&
        memcmp(&g.out, out, sizeof(*out)) == 0) {
      ISV_CUDA(cudaGraphLaunch(g.exec, s));
      h->launches += g.launches;
      h->bgraph_miss = 0;
      return ISV_OK;
    }
  }
  // Measured on B200 (L = 1000 / 150): replay wins while the host's issue latency is on the critical path -- one window
  // 51 -> 42 us, 512 windows 73 -> 65 us, 4096 windows 148 -> 141 us -- and loses 2.5 % at 9472 windows (0.385 -> 0.395
  // ms), where the stream-ordered launches already overlap the three chains better than the graph's schedule does.
  constexpr int kGraphMaxWindows = 4736;   // two waves of 148 SMs x 16 resident windows
  static const bool no_graph = getenv("ISV_NO_GRAPH") != nullptr;   // A/B switch for measurements
  if (no_graph || in->n_windows > kGraphMaxWindows) return launch_batch(h, in, out, which, s, gram);
  if (++h->bgraph_miss > 8) return launch_batch(h, in, out, which, s, gram);
  const int64_t l0 = h->launches;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
    const isv_status stc = launch_batch(h, in, out, which, s, gram);
    const cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (stc == ISV_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
      cudaGraphDestroy(graph);
      isv_handle::BatchGraph& g = h->bgraph[h->bgraph_next];
      h->bgraph_next = (h->bgraph_next + 1) & 3;
      if (g.exec) cudaGraphExecDestroy(g.exec);
      g.in = *in; g.out = *out; g.which = which; g.stream = s; g.gram = gram; g.exec = exec; g.launches = (int)(h->launches - l0);
      ISV_CUDA(cudaGraphLaunch(exec, s));
      return ISV_OK;
    }
    if (graph) cudaGraphDestroy(graph);
  }
  cudaGetLastError();      // capture not possible (e.g. the caller's stream is already capturing): issue directly
  h->launches = l0;
  return launch_batch(h, in, out, which, s, gram);
}

// ---- host-pointer entry point ------------------------------------------------------------------
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static isv_status ensure_dbuf(isv_handle* h, size_t bytes) {
  if (h->dbuf_bytes >= bytes) return ISV_OK;
  if (h->dbuf) {
    ISV_CUDA(cudaStreamSynchronize(h->stream));
    ISV_CUDA(cudaFree(h->dbuf));
    h->dbuf = nullptr;
    h->dbuf_bytes = 0;
  }
  size_t cap = bytes + bytes / 4;
  if (cudaMalloc(&h->dbuf, cap) != cudaSuccess) {
    cudaGetLastError();
    return ISV_ERR_ALLOC;
  }
  h->dbuf_bytes = cap;
  return ISV_OK;
}

isv_status isv_marg_window_batch_host(isv_handle* h, const isv_batch_in* in, const isv_batch_out* out, int which) {
  if (!h) return ISV_ERR_BAD_ARG;
  if (which & ~ISV_RUN_BOTH) return ISV_ERR_BAD_ARG;   // the stage flags are device-path profiling aids
  isv_status st = check_batch(in, out, which, true);
  if (st != ISV_OK) return st;
  const size_t n = (size_t)in->n_windows;
  if (n == 0) return ISV_OK;
  ISV_CUDA(cudaSetDevice(h->device));
  const bool fwd = which & ISV_RUN_FORWARD, bwd = which & ISV_RUN_BACKWARD;
  const size_t D = sizeof(double);
  const int64_t n_lm = fwd ? in->lm_offset[n] : 0;
  if (n_lm < 0 || (fwd && in->lm_stride < n_lm)) return ISV_ERR_BAD_ARG;
  // carve the device mirror (whole batch; chunks address sub-ranges of it)
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  const size_t o_lmoff = carve(fwd ? (n + 1) * sizeof(int64_t) : 0);
  const size_t o_obs = carve(fwd ? 6 * (size_t)n_lm * D : 0);
  const bool xyf = fwd && in->lm_xy_f32;
  const size_t o_xyf = carve(xyf ? 2 * (size_t)n_lm * sizeof(float) : 0);
  const size_t o_posef = carve(fwd ? n * 14 * D : 0);
  const size_t o_ex = carve(fwd ? (in->ex_pose_shared ? 7 : n * 7) * D : 0);
  const size_t o_pse3 = carve(fwd ? n * ISV_SE3_REC * D : 0);
  const size_t o_prel = carve(fwd ? n * ISV_REL_REC * D : 0);
  const size_t o_prp = carve(fwd && in->prior_rp ? n * ISV_RP_IN_REC * D : 0);
  const size_t o_poseb = carve(bwd ? n * 14 * D : 0);
  const size_t o_sbb = carve(bwd ? n * 18 * D : 0);
  const size_t o_pvb = carve(bwd ? n * ISV_VB_REC * D : 0);
  const bool raw_imu = bwd && !in->preint;      // ABI 2: ship the raw samples, rebuild the record on the GPU
  const size_t K = raw_imu ? (size_t)in->imu_k_max : 0;
  const size_t o_pre = carve(bwd && !raw_imu ? n * ISV_PREINT_REC * D : 0);
  const size_t o_iraw = carve(raw_imu ? n * K * 7 * D : 0);
  const size_t o_iinit = carve(raw_imu ? n * 12 * D : 0);
  const size_t o_icnt = carve(raw_imu && in->imu_count ? n * sizeof(int32_t) : 0);
  const bool z_one = fwd && (in->flags & ISV_IN_PTS_I_Z_ONE);
  if (z_one && n_lm > 0) {   // spot-check the caller's promise: first, last and every 4096th landmark
    const double* z = in->lm_obs + 2 * (size_t)in->lm_stride;
    for (int64_t k = 0; k < n_lm; k += 4096)
      if (z[k] != 1.0) return ISV_ERR_BAD_ARG;
    if (z[n_lm - 1] != 1.0) return ISV_ERR_BAD_ARG;
  }
  if (raw_imu && in->imu_count)
    for (size_t w = 0; w < n; ++w)
      if (in->imu_count[w] < 0 || in->imu_count[w] > in->imu_k_max) return ISV_ERR_BAD_ARG;
  const size_t o_se3 = carve(fwd ? n * ISV_SE3_REC * D : 0);
  const size_t o_pg = carve(fwd ? n * ISV_PG_REC * D : 0);
  const size_t o_rel = carve(bwd ? n * ISV_REL_REC * D : 0);
  const size_t o_vb = carve(bwd ? n * ISV_VB_REC * D : 0);
  const size_t o_rp = carve(bwd ? n * ISV_RP_REC * D : 0);
  const size_t o_rank = carve(n * 2 * sizeof(int32_t));
  const size_t o_stat = carve(n * sizeof(int32_t));
  const size_t o_gram = carve(n * kScratchPerWindow * D);   // per-chunk slices: [w0 * kScratchPerWindow ...)
  const size_t o_gramB = carve(bwd ? n * kScratchPerWindow * D : 0);   // the whole-batch backward call's own scratch
  const bool tri_in = (in->flags & ISV_IN_TRI_RECORDS) != 0, tri_out = (in->flags & ISV_OUT_TRI_RECORDS) != 0;
  const size_t p_se3 = carve(tri_in && fwd ? n * ISV_SE3_TRI_REC * D : 0), p_rel = carve(tri_in && fwd ? n * ISV_REL_TRI_REC * D : 0);
  const size_t p_rp = carve(tri_in && fwd && in->prior_rp ? n * ISV_RP_IN_TRI_REC * D : 0), p_vb = carve(tri_in && bwd ? n * ISV_VB_TRI_REC * D : 0);
  const size_t q_se3 = carve(tri_out && fwd ? n * ISV_SE3_TRI_REC * D : 0), q_pg = carve(tri_out && fwd ? n * ISV_PG_TRI_REC * D : 0);
  const size_t q_rel = carve(tri_out && bwd ? n * ISV_REL_TRI_REC * D : 0), q_vb = carve(tri_out && bwd ? n * ISV_VB_TRI_REC * D : 0);
  const size_t q_rp = carve(tri_out && bwd ? n * ISV_RP_TRI_REC * D : 0);
  const int kflags = in->flags & ISV_IN_PTS_I_Z_ONE;   // what the kernels know about
  st = ensure_dbuf(h, off);
  if (st != ISV_OK) return st;
  char* d = h->dbuf;
  // Chunked pipeline on two streams: chunk c+1's H2D overlaps chunk c's kernels and D2H (PCIe is
  // full duplex).  The caller's stream is fenced before and after with events.
  // Four pipeline streams: every chunk is a dependent chain H2D -> kernels -> D2H, and at L ~ 150 that chain is
  // latency- not bandwidth-bound (0.12 + 0.39 + 0.25 ms per 2368 windows, profiles/r02t_host_pipeline_trace.txt): with two
  // streams only two chunks were ever in flight.
  constexpr int kPipe = 4;
  cudaStream_t ss[kPipe] = {h->own_stream, h->copy_stream, h->pipe[0], h->pipe[1]};
  // All H2D copies go through ONE stream, in chunk order (the copy engine serves them at line rate either way, but spread
  // over the four pipeline streams it picked its own order -- chunk 3 landed before chunks 1 and 2 -- so the chunk that
  // arrives last, whose kernels and D2H nothing overlaps, could not be chosen); chunk c's pipeline stream waits for
  // "chunk c is on the device".
  cudaStream_t hs = h->h2d_stream;
  if (h->stream != h->own_stream) {
    ISV_CUDA(cudaEventRecord(h->ev[0], h->stream));
    for (int i = 0; i < kPipe; ++i) ISV_CUDA(cudaStreamWaitEvent(ss[i], h->ev[0], 0));
  } else {
    ISV_CUDA(cudaEventRecord(h->ev[0], ss[0]));
    for (int i = 1; i < kPipe; ++i) ISV_CUDA(cudaStreamWaitEvent(ss[i], h->ev[0], 0));
  }
  ISV_CUDA(cudaStreamWaitEvent(hs, h->ev[0], 0));
  // rank and status are zeroed once, here: the forward chunks and the whole-batch backward call OR into the same status words
  zero_i32_kernel<<<(int)((2 * n + 255) / 256), 256, 0, ss[0]>>>((int32_t*)(d + o_rank), (long long)(2 * n));
  zero_i32_kernel<<<(int)((n + 255) / 256), 256, 0, ss[0]>>>((int32_t*)(d + o_stat), (long long)n);
  // The per-window records cross once, one copy per array for the whole batch, ahead of the chunk loop: ~320 doubles per
  // window that used to travel as nine copies PER CHUNK (every copy pays a fixed DMA set-up; at L = 150 they are half of the
  // bytes).  Only the landmark components and the results are chunked.
  auto up = [&](size_t o, const void* src, size_t bytes) {
    return bytes ? cudaMemcpyAsync(d + o, src, bytes, cudaMemcpyHostToDevice, hs) : cudaSuccess;
  };
  if (fwd) {
    ISV_CUDA(up(o_lmoff, in->lm_offset, (n + 1) * sizeof(int64_t)));
    ISV_CUDA(up(o_ex, in->ex_pose, (in->ex_pose_shared ? 7 : n * 7) * D));
    ISV_CUDA(up(o_posef, in->pose_fwd, n * 14 * D));
    if (!tri_in) {
      ISV_CUDA(up(o_pse3, in->prior_se3, n * ISV_SE3_REC * D));
      ISV_CUDA(up(o_prel, in->prior_rel, n * ISV_REL_REC * D));
      if (in->prior_rp) ISV_CUDA(up(o_prp, in->prior_rp, n * ISV_RP_IN_REC * D));
    } else {   // ABI 4: packed records up, expanded on the device (on the H2D stream: a few microseconds between two copies)
      ISV_CUDA(up(p_se3, in->prior_se3, n * ISV_SE3_TRI_REC * D));
      ISV_CUDA(up(p_rel, in->prior_rel, n * ISV_REL_TRI_REC * D));
      if (in->prior_rp) ISV_CUDA(up(p_rp, in->prior_rp, n * ISV_RP_IN_TRI_REC * D));
      tri_launch<false>(h, hs, n, (const double*)(d + p_se3), (double*)(d + o_pse3), kTriSe3);
      tri_launch<false>(h, hs, n, (const double*)(d + p_rel), (double*)(d + o_prel), kTriSe3);
      if (in->prior_rp) tri_launch<false>(h, hs, n, (const double*)(d + p_rp), (double*)(d + o_prp), kTriRpIn);
    }
  }
  ISV_CUDA(cudaEventRecord(h->ev[1], ss[0]));      // the rank / status zero-fill
  for (int i = 1; i < kPipe; ++i) ISV_CUDA(cudaStreamWaitEvent(ss[i], h->ev[1], 0));
  // ---- the backward half: MargBackward reads no landmark, only the per-window records that have just been queued -- it
  // runs ONCE for the whole batch (whole waves instead of four latency-bound partial ones, a quarter of the launches) on
  // its own stream, under the landmark H2D, and its three result arrays go back as three copies.  Only the forward half
  // (landmark phase + tail) is chunked behind the landmark stream.
  cudaStream_t sb = h->bwd_stream;
  // (When only the backward half is wanted nothing hides it, so it is pipelined itself: four sub-batches, each with its own
  // record copies, on the four pipeline streams.)
  const size_t nb = bwd ? ((!fwd && n >= 2048) ? 4 : 1) : 0;
  for (size_t k = 0; k < nb; ++k) {
    const size_t w0 = n * k / nb, w1 = n * (k + 1) / nb, m = w1 - w0;
    if (m == 0) continue;
    cudaStream_t sk = nb == 1 ? sb : ss[k % kPipe];
    auto upr = [&](size_t o, const void* src, size_t rec_bytes) {   // rows [w0, w1) of a per-window array
      return cudaMemcpyAsync(d + o + w0 * rec_bytes, (const char*)src + w0 * rec_bytes, m * rec_bytes, cudaMemcpyHostToDevice, hs);
    };
    ISV_CUDA(upr(o_poseb, in->pose_bwd, 14 * D));
    ISV_CUDA(upr(o_sbb, in->sb_bwd, 18 * D));
    if (!tri_in) {
      ISV_CUDA(upr(o_pvb, in->prior_vb, ISV_VB_REC * D));
    } else {
      ISV_CUDA(upr(p_vb, in->prior_vb, ISV_VB_TRI_REC * D));
      tri_launch<false>(h, hs, m, (const double*)(d + p_vb) + w0 * ISV_VB_TRI_REC, (double*)(d + o_pvb) + w0 * ISV_VB_REC, kTriVb);
    }
    if (!raw_imu) {
      ISV_CUDA(upr(o_pre, in->preint, ISV_PREINT_REC * D));
    } else {
      if (K) ISV_CUDA(upr(o_iraw, in->imu_raw, K * 7 * D));
      ISV_CUDA(upr(o_iinit, in->imu_init, 12 * D));
      if (in->imu_count) ISV_CUDA(upr(o_icnt, in->imu_count, sizeof(int32_t)));
    }
    cudaEvent_t ev_rec = h->chunk_ev[12 + k];      // "the records of sub-batch k are on the device"
    ISV_CUDA(cudaEventRecord(ev_rec, hs));
    ISV_CUDA(cudaStreamWaitEvent(sk, h->ev[1], 0));
    ISV_CUDA(cudaStreamWaitEvent(sk, ev_rec, 0));
    isv_batch_in bin;
    isv_batch_out bout;
    memset(&bin, 0, sizeof(bin));
    memset(&bout, 0, sizeof(bout));
    bin.n_windows = (int32_t)m;
    bin.flags = kflags;
    bin.pose_bwd = (const double*)(d + o_poseb) + w0 * 14;
    bin.sb_bwd = (const double*)(d + o_sbb) + w0 * 18;
    bin.prior_vb = (const double*)(d + o_pvb) + w0 * ISV_VB_REC;
    if (raw_imu) {
      bin.imu_raw = (const double*)(d + o_iraw) + w0 * K * 7;
      bin.imu_init = (const double*)(d + o_iinit) + w0 * 12;
      bin.imu_count = in->imu_count ? (const int32_t*)(d + o_icnt) + w0 : nullptr;
      bin.imu_k_max = in->imu_k_max;
    } else {
      bin.preint = (const double*)(d + o_pre) + w0 * ISV_PREINT_REC;
    }
    bout.rel_out = (double*)(d + o_rel) + w0 * ISV_REL_REC;
    bout.vb_out = (double*)(d + o_vb) + w0 * ISV_VB_REC;
    bout.rp_out = (double*)(d + o_rp) + w0 * ISV_RP_REC;
    bout.rank = (int32_t*)(d + o_rank) + 2 * w0;
    bout.status = (int32_t*)(d + o_stat) + w0;
    st = launch_batch(h, &bin, &bout, ISV_RUN_BACKWARD, sk, (double*)(d + o_gramB) + w0 * kScratchPerWindow, DbgStores{nullptr, nullptr},
                      false);
    if (st != ISV_OK) return st;
    if (nb == 1) ISV_CUDA(cudaEventRecord(h->ev[3], sk));   // backward kernels done: rank[.][1] and its status bits are final
    if (!tri_out) {
      ISV_CUDA(cudaMemcpyAsync(out->rel_out + w0 * ISV_REL_REC, bout.rel_out, m * ISV_REL_REC * D, cudaMemcpyDeviceToHost, sk));
      ISV_CUDA(cudaMemcpyAsync(out->vb_out + w0 * ISV_VB_REC, bout.vb_out, m * ISV_VB_REC * D, cudaMemcpyDeviceToHost, sk));
      ISV_CUDA(cudaMemcpyAsync(out->rp_out + w0 * ISV_RP_REC, bout.rp_out, m * ISV_RP_REC * D, cudaMemcpyDeviceToHost, sk));
    } else {   // ABI 4: compacted on the device, 99 instead of 151 doubles per window on the way back
      double* qr = (double*)(d + q_rel) + w0 * ISV_REL_TRI_REC;
      double* qv = (double*)(d + q_vb) + w0 * ISV_VB_TRI_REC;
      double* qp = (double*)(d + q_rp) + w0 * ISV_RP_TRI_REC;
      tri_launch<true>(h, sk, m, bout.rel_out, qr, kTriSe3);
      tri_launch<true>(h, sk, m, bout.vb_out, qv, kTriVb);
      tri_launch<true>(h, sk, m, bout.rp_out, qp, kTriRpOut);
      ISV_CUDA(cudaMemcpyAsync(out->rel_out + w0 * ISV_REL_TRI_REC, qr, m * ISV_REL_TRI_REC * D, cudaMemcpyDeviceToHost, sk));
      ISV_CUDA(cudaMemcpyAsync(out->vb_out + w0 * ISV_VB_TRI_REC, qv, m * ISV_VB_TRI_REC * D, cudaMemcpyDeviceToHost, sk));
      ISV_CUDA(cudaMemcpyAsync(out->rp_out + w0 * ISV_RP_TRI_REC, qp, m * ISV_RP_TRI_REC * D, cudaMemcpyDeviceToHost, sk));
    }
    if (!fwd) {
      ISV_CUDA(cudaMemcpyAsync(out->rank + 2 * w0, bout.rank, m * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, sk));
      if (out->status) ISV_CUDA(cudaMemcpyAsync(out->status + w0, bout.status, m * sizeof(int32_t), cudaMemcpyDeviceToHost, sk));
    }
  }
  size_t n_chunks = n / 512;
  if (n_chunks < 1) n_chunks = 1;
  static const int max_chunks = getenv("ISV_HOST_CHUNKS") ? atoi(getenv("ISV_HOST_CHUNKS")) : 4;   // measured at 9472 windows: 2 / 4 / 8 / 16 chunks
                                                                                                   // -> 1.68 / 1.76 / 1.73 / 1.56 M windows/s
  if (n_chunks > (size_t)max_chunks) n_chunks = max_chunks > 0 ? max_chunks : 1;
  if (n_chunks > 16) n_chunks = 16;
  // The call ends with the LAST chunk's kernels and D2H, which nothing overlaps: when the landmark stream dominates the
  // bytes (L in the hundreds and up) the last of four chunks is the smallest (30 / 30 / 25 / 15 % of the windows: 3.71 ->
  // 3.66 ms at L ~ 1000; at L ~ 150 the chunks are latency-bound and equal sizes are better, 1.47 vs 1.51 ms).
  // ISV_HOST_EVEN_CHUNKS=1 forces equal chunks for A/B measurements.
  static const bool even_env = getenv("ISV_HOST_EVEN_CHUNKS") != nullptr;
  const bool even_chunks = even_env || (size_t)n_lm * 16 < (size_t)n * 8192;
  auto bound = [&](size_t c) -> size_t {
    if (n_chunks == 4 && !even_chunks && n >= 64) {
      static const double cum[5] = {0.0, 0.30, 0.60, 0.85, 1.0};
      return c >= 4 ? n : (size_t)(cum[c] * (double)n);
    }
    return n * c / n_chunks;
  };
  // ISV_HOST_TRACE=1: timeline of the chunk pipeline (device events + host issue times), printed to stderr
  static const bool trace = getenv("ISV_HOST_TRACE") != nullptr;
  cudaEvent_t tev[1 + 4 * 16];
  double thost[16][4];
  timespec tr0 = {};
  const bool tr = trace && n_chunks <= 16;
  auto host_ms = [&]() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (t.tv_sec - tr0.tv_sec) * 1e3 + (t.tv_nsec - tr0.tv_nsec) * 1e-6; };
  if (tr) {
    for (size_t i = 0; i < 1 + 4 * n_chunks; ++i) cudaEventCreate(&tev[i]);
    clock_gettime(CLOCK_MONOTONIC, &tr0);
    cudaEventRecord(tev[0], ss[0]);
  }
  for (size_t c = 0; fwd && c < n_chunks; ++c) {
    const size_t w0 = bound(c), w1 = bound(c + 1), m = w1 - w0;
    if (m == 0) continue;
    cudaStream_t s = ss[c % kPipe];
    if (tr) { thost[c][0] = host_ms(); cudaEventRecord(tev[1 + 4 * c], s); }
    isv_batch_in din;
    isv_batch_out dout;
    memset(&din, 0, sizeof(din));
    memset(&dout, 0, sizeof(dout));
    din.n_windows = (int32_t)m;
    din.ex_pose_shared = in->ex_pose_shared;
    din.flags = kflags;
    if (fwd) {
      const int64_t a = in->lm_offset[w0], b = in->lm_offset[w1];
      // components 3,4 (pts_j) are never read by the information-only marginalization: not copied
      static const int comps[4] = {5, 2, 0, 1};     // inv_dep; pts_i.z (skipped under ISV_IN_PTS_I_Z_ONE); pts_i.x, pts_i.y
      for (int ci = 0; ci < 4 && b > a; ++ci) {
        if ((comps[ci] == 2 && z_one) || (comps[ci] < 2 && xyf)) continue;
        ISV_CUDA(cudaMemcpyAsync(d + o_obs + ((size_t)comps[ci] * n_lm + a) * D,
                                 in->lm_obs + (size_t)comps[ci] * in->lm_stride + a, (size_t)(b - a) * D,
                                 cudaMemcpyHostToDevice, hs));
      }
      if (xyf && b > a)   // ABI 3: the FP32 x, y of the feature tracker, 8 instead of 16 bytes per landmark
        for (int ci = 0; ci < 2; ++ci)
          ISV_CUDA(cudaMemcpyAsync(d + o_xyf + ((size_t)ci * n_lm + a) * sizeof(float),
                                   in->lm_xy_f32 + (size_t)ci * in->lm_stride + a, (size_t)(b - a) * sizeof(float),
                                   cudaMemcpyHostToDevice, hs));
      din.lm_offset = (const int64_t*)(d + o_lmoff) + w0;   // absolute offsets into the whole mirror
      din.lm_obs = (const double*)(d + o_obs);
      din.lm_stride = n_lm;
      din.lm_xy_f32 = xyf ? (const float*)(d + o_xyf) : nullptr;
      din.pose_fwd = (const double*)(d + o_posef) + w0 * 14;
      din.ex_pose = (const double*)(d + o_ex) + (in->ex_pose_shared ? 0 : w0 * 7);
      din.prior_se3 = (const double*)(d + o_pse3) + w0 * ISV_SE3_REC;
      din.prior_rel = (const double*)(d + o_prel) + w0 * ISV_REL_REC;
      din.prior_rp = in->prior_rp ? (const double*)(d + o_prp) + w0 * ISV_RP_IN_REC : nullptr;
      dout.se3_out = (double*)(d + o_se3) + w0 * ISV_SE3_REC;
      dout.pg_out = (double*)(d + o_pg) + w0 * ISV_PG_REC;
    }
    dout.rank = (int32_t*)(d + o_rank) + 2 * w0;
    dout.status = (int32_t*)(d + o_stat) + w0;
    ISV_CUDA(cudaEventRecord(h->chunk_ev[c], hs));          // records + chunks 0 .. c are on the device
    ISV_CUDA(cudaStreamWaitEvent(s, h->chunk_ev[c], 0));
    if (tr) { thost[c][1] = host_ms(); cudaEventRecord(tev[2 + 4 * c], s); }
    st = launch_batch(h, &din, &dout, ISV_RUN_FORWARD, s, (double*)(d + o_gram) + w0 * kScratchPerWindow, DbgStores{nullptr, nullptr},
                      false);
    if (st != ISV_OK) return st;
    if (tr) { thost[c][2] = host_ms(); cudaEventRecord(tev[3 + 4 * c], s); }
    if (fwd && !tri_out) {
      ISV_CUDA(cudaMemcpyAsync(out->se3_out + w0 * ISV_SE3_REC, dout.se3_out, m * ISV_SE3_REC * D, cudaMemcpyDeviceToHost, s));
      ISV_CUDA(cudaMemcpyAsync(out->pg_out + w0 * ISV_PG_REC, dout.pg_out, m * ISV_PG_REC * D, cudaMemcpyDeviceToHost, s));
    } else if (fwd) {
      double* qs = (double*)(d + q_se3) + w0 * ISV_SE3_TRI_REC;
      double* qp = (double*)(d + q_pg) + w0 * ISV_PG_TRI_REC;
      tri_launch<true>(h, s, m, dout.se3_out, qs, kTriSe3);
      tri_launch<true>(h, s, m, dout.pg_out, qp, kTriPg);
      ISV_CUDA(cudaMemcpyAsync(out->se3_out + w0 * ISV_SE3_TRI_REC, qs, m * ISV_SE3_TRI_REC * D, cudaMemcpyDeviceToHost, s));
      ISV_CUDA(cudaMemcpyAsync(out->pg_out + w0 * ISV_PG_TRI_REC, qp, m * ISV_PG_TRI_REC * D, cudaMemcpyDeviceToHost, s));
    }
    if (bwd) ISV_CUDA(cudaStreamWaitEvent(s, h->ev[3], 0));   // rank / status of the chunk's windows carry the backward half too
    ISV_CUDA(cudaMemcpyAsync(out->rank + 2 * w0, dout.rank, m * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (out->status)
      ISV_CUDA(cudaMemcpyAsync(out->status + w0, dout.status, m * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (tr) { thost[c][3] = host_ms(); cudaEventRecord(tev[4 + 4 * c], s); }
  }
  for (int i = 0; i < kPipe; ++i) ISV_CUDA(cudaStreamSynchronize(ss[i]));
  ISV_CUDA(cudaStreamSynchronize(hs));
  if (bwd) ISV_CUDA(cudaStreamSynchronize(sb));
  if (tr) {
    const double t_end = host_ms();
    for (size_t c = 0; c < n_chunks; ++c) {
      float e[4];
      for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&e[k], tev[0], tev[1 + 4 * c + k]);
      fprintf(stderr, "[isv trace] chunk %zu: device ms  start %.3f | H2D done %.3f | kernels done %.3f | D2H done %.3f   ||  host issue ms  %.3f %.3f %.3f %.3f\n",
              c, e[0], e[1], e[2], e[3], thost[c][0], thost[c][1], thost[c][2], thost[c][3]);
    }
    fprintf(stderr, "[isv trace] host: call returned at %.3f ms\n", t_end);
    for (size_t i = 0; i < 1 + 4 * n_chunks; ++i) cudaEventDestroy(tev[i]);
  }
  return ISV_OK;
}

// ---- forensic mode: the structured path's intermediates, the KLD diagnostics, the reference's dense Schur route ------
isv_status isv_marg_forensic_batch(isv_handle* h, const isv_batch_in* in, const isv_batch_out* out, const isv_forensic_out* dbg) {
  if (!h || !dbg) return ISV_ERR_BAD_ARG;
  isv_status st = check_batch(in, out, ISV_RUN_BOTH);
  if (st != ISV_OK) return st;
  if (!dbg->lamda_prior_fwd || !dbg->g_bwd || !dbg->kld_fwd || !dbg->kld_bwd) return ISV_ERR_BAD_ARG;
  const int n = in->n_windows;
  if (n == 0) return ISV_OK;
  ISV_CUDA(cudaSetDevice(h->device));
  const size_t need = (size_t)n * kScratchPerWindow * sizeof(double);
  if (h->gram_bytes < need) {
    if (h->gram) { ISV_CUDA(cudaStreamSynchronize(h->stream)); ISV_CUDA(cudaFree(h->gram)); h->gram = nullptr; h->gram_bytes = 0; }
    if (cudaMalloc(&h->gram, need + 256) != cudaSuccess) { cudaGetLastError(); return ISV_ERR_ALLOC; }
    h->gram_bytes = need + 256;
  }
  st = launch_batch(h, in, out, ISV_RUN_BOTH, h->stream, h->gram, DbgStores{dbg->lamda_prior_fwd, dbg->g_bwd});
  if (st != ISV_OK) return st;
  st = ensure_dbuf(h, (size_t)n * kKldScratch * sizeof(double));   // the host-pointer staging buffer is idle on this path
  if (st != ISV_OK) return st;
  isv_kld_args a;
  a.scratch = (double*)h->dbuf;
  a.n = n;
  a.pose_bwd = in->pose_bwd;
  a.fj = h->gram + (size_t)n * 42;
  a.lp_fwd = dbg->lamda_prior_fwd;
  a.g_bwd = dbg->g_bwd;
  a.se3_out = out->se3_out; a.rel_out = out->rel_out; a.vb_out = out->vb_out; a.rp_out = out->rp_out;
  a.rank = out->rank;
  a.alpha = h->dcfg.alpha;
  a.kld_fwd = dbg->kld_fwd; a.kld_bwd = dbg->kld_bwd;
  a.lp_bwd = dbg->lamda_prior_bwd; a.eig_bwd = dbg->eig_bwd; a.info_abs = dbg->info_abs; a.info_yaw = dbg->info_yaw;
  marg_kld_kernel<<<(n + 63) / 64, 64, 0, h->stream>>>(a);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

isv_status isv_literal_schur(isv_handle* h, int n_problems, int n, int m0, const double* A, double* A_prior, double* Amm_inv,
                             int32_t* rank) {
  if (!h || n_problems < 1 || m0 < 1 || n <= m0 || n - m0 > 1024 || !A || !A_prior) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  const size_t m = (size_t)(n - m0);
  isv_status st = ensure_dbuf(h, (size_t)n_problems * m * 2 * m * sizeof(double));
  if (st != ISV_OK) return st;
  literal_schur_kernel<<<n_problems, kLitThreads, m * sizeof(int), h->stream>>>(n, m0, A, (double*)h->dbuf, A_prior, Amm_inv, rank);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

// ---- single-window wrappers --------------------------------------------------------------------
// One window, latency path: every input is packed into ONE pinned staging block (one H2D), the outputs
// come back as ONE block (one D2H); ~2.5x lower latency than going through the chunked batch pipeline.
static isv_status ensure_pinned(isv_handle* h, size_t bytes) {
  if (h->pinned_bytes >= bytes) return ISV_OK;
  if (h->pinned) {
    cudaStreamSynchronize(h->stream);
    cudaFreeHost(h->pinned);
    h->pinned = nullptr;
    h->pinned_bytes = 0;
  }
  if (cudaMallocHost(&h->pinned, bytes + bytes / 2) != cudaSuccess) {
    cudaGetLastError();
    return ISV_ERR_ALLOC;
  }
  h->pinned_bytes = bytes + bytes / 2;
  return ISV_OK;
}

isv_status isv_marg_forward(isv_handle* h, const isv_fwd_in* in, isv_fwd_out* out) {
  if (!h || !in || !out || in->n_landmarks < 0) return ISV_ERR_BAD_ARG;
  if (!in->pose0 || !in->pose1 || !in->ex_pose || !in->prior_se3 || !in->prior_rel) return ISV_ERR_BAD_ARG;
  const size_t L = (size_t)in->n_landmarks;
  if (L > 0 && (!in->inv_dep || !in->pts_i || !in->pts_j)) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  // block layout (doubles): in  = [lm_offset 2 (as int64)] [obs 6L] [pose_fwd 14] [ex 7] [se3 48] [rel 48] [rp 5]
  //                          out = [se3 48] [pg 89] [rank 2 x i32 = 1] [status i32 = 1]
  const size_t n_in = 2 + 6 * L + 14 + 7 + ISV_SE3_REC + ISV_REL_REC + ISV_RP_IN_REC;
  const size_t n_out = ISV_SE3_REC + ISV_PG_REC + 2;
  const size_t n_scr = kScratchPerWindow;
  isv_status st = ensure_pinned(h, (n_in + n_out) * sizeof(double));
  if (st != ISV_OK) return st;
  st = ensure_dbuf(h, (n_in + n_out + n_scr + 64) * sizeof(double));
  if (st != ISV_OK) return st;
  double* hp = (double*)h->pinned;
  int64_t* off = (int64_t*)hp;
  off[0] = 0; off[1] = (int64_t)L;
  double* obs = hp + 2;
  for (size_t k = 0; k < L; ++k) {
    obs[k] = in->pts_i[3 * k];
    obs[L + k] = in->pts_i[3 * k + 1];
    obs[2 * L + k] = in->pts_i[3 * k + 2];
    obs[3 * L + k] = in->pts_j[3 * k];
    obs[4 * L + k] = in->pts_j[3 * k + 1];
    obs[5 * L + k] = in->inv_dep[k];
  }
  double* q = obs + 6 * L;
  memcpy(q, in->pose0, 56); memcpy(q + 7, in->pose1, 56); memcpy(q + 14, in->ex_pose, 56);
  memcpy(q + 21, in->prior_se3, ISV_SE3_REC * 8); memcpy(q + 21 + ISV_SE3_REC, in->prior_rel, ISV_REL_REC * 8);
  double* rp = q + 21 + ISV_SE3_REC + ISV_REL_REC;
  if (in->prior_rp) memcpy(rp, in->prior_rp, ISV_RP_IN_REC * 8); else memset(rp, 0, ISV_RP_IN_REC * 8);
  double* d = (double*)h->dbuf;
  cudaStream_t s = h->stream;
  ISV_CUDA(cudaMemcpyAsync(d, hp, n_in * sizeof(double), cudaMemcpyHostToDevice, s));
  double* dq = d + 2 + 6 * L;
  double* dout = d + n_in;
  isv_batch_in bi;
  memset(&bi, 0, sizeof(bi));
  bi.n_windows = 1;
  bi.ex_pose_shared = 1;
  bi.lm_offset = (const int64_t*)d;
  bi.lm_obs = d + 2;
  bi.lm_stride = (int64_t)L;
  bi.pose_fwd = dq;
  bi.ex_pose = dq + 14;
  bi.prior_se3 = dq + 21;
  bi.prior_rel = dq + 21 + ISV_SE3_REC;
  bi.prior_rp = dq + 21 + ISV_SE3_REC + ISV_REL_REC;
  isv_batch_out bo;
  memset(&bo, 0, sizeof(bo));
  bo.se3_out = dout;
  bo.pg_out = dout + ISV_SE3_REC;
  bo.rank = (int32_t*)(dout + ISV_SE3_REC + ISV_PG_REC);
  bo.status = (int32_t*)(dout + ISV_SE3_REC + ISV_PG_REC + 1);
  ISV_CUDA(cudaMemsetAsync(bo.rank, 0, 8, s));
  st = launch_batch(h, &bi, &bo, ISV_RUN_FORWARD, s, dout + n_out);
  if (st != ISV_OK) return st;
  double* ho = hp + n_in;
  ISV_CUDA(cudaMemcpyAsync(ho, dout, n_out * sizeof(double), cudaMemcpyDeviceToHost, s));
  ISV_CUDA(cudaStreamSynchronize(s));
  memcpy(out->se3, ho, ISV_SE3_REC * 8);
  memcpy(out->pg, ho + ISV_SE3_REC, ISV_PG_REC * 8);
  out->rank = ((const int32_t*)(ho + ISV_SE3_REC + ISV_PG_REC))[0];
  out->status = ((const int32_t*)(ho + ISV_SE3_REC + ISV_PG_REC + 1))[0];
  return ISV_OK;
}

isv_status isv_marg_backward(isv_handle* h, const isv_bwd_in* in, isv_bwd_out* out) {
  if (!h || !in || !out) return ISV_ERR_BAD_ARG;
  if (!in->pose_i || !in->sb_i || !in->pose_j || !in->sb_j || !in->prior_vb || !in->preint) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  // in = [pose_bwd 14] [sb_bwd 18] [vb 90] [preint 467] ; out = [rel 48] [vb 90] [rp 13] [rank 1] [status 1]
  const size_t n_in = 14 + 18 + ISV_VB_REC + ISV_PREINT_REC;
  const size_t n_out = ISV_REL_REC + ISV_VB_REC + ISV_RP_REC + 2;
  isv_status st = ensure_pinned(h, (n_in + n_out) * sizeof(double));
  if (st != ISV_OK) return st;
  st = ensure_dbuf(h, (n_in + n_out + kScratchPerWindow + 64) * sizeof(double));
  if (st != ISV_OK) return st;
  double* hp = (double*)h->pinned;
  memcpy(hp, in->pose_i, 56); memcpy(hp + 7, in->pose_j, 56);
  memcpy(hp + 14, in->sb_i, 72); memcpy(hp + 23, in->sb_j, 72);
  memcpy(hp + 32, in->prior_vb, ISV_VB_REC * 8);
  memcpy(hp + 32 + ISV_VB_REC, in->preint, ISV_PREINT_REC * 8);
  double* d = (double*)h->dbuf;
  cudaStream_t s = h->stream;
  ISV_CUDA(cudaMemcpyAsync(d, hp, n_in * sizeof(double), cudaMemcpyHostToDevice, s));
  double* dout = d + n_in;
  isv_batch_in bi;
  memset(&bi, 0, sizeof(bi));
  bi.n_windows = 1;
  bi.pose_bwd = d;
  bi.sb_bwd = d + 14;
  bi.prior_vb = d + 32;
  bi.preint = d + 32 + ISV_VB_REC;
  isv_batch_out bo;
  memset(&bo, 0, sizeof(bo));
  bo.rel_out = dout;
  bo.vb_out = dout + ISV_REL_REC;
  bo.rp_out = dout + ISV_REL_REC + ISV_VB_REC;
  bo.rank = (int32_t*)(dout + ISV_REL_REC + ISV_VB_REC + ISV_RP_REC);
  bo.status = (int32_t*)(dout + ISV_REL_REC + ISV_VB_REC + ISV_RP_REC + 1);
  ISV_CUDA(cudaMemsetAsync(bo.rank, 0, 8, s));
  st = launch_batch(h, &bi, &bo, ISV_RUN_BACKWARD, s, dout + n_out);
  if (st != ISV_OK) return st;
  double* ho = hp + n_in;
  ISV_CUDA(cudaMemcpyAsync(ho, dout, n_out * sizeof(double), cudaMemcpyDeviceToHost, s));
  ISV_CUDA(cudaStreamSynchronize(s));
  memcpy(out->rel, ho, ISV_REL_REC * 8);
  memcpy(out->vb, ho + ISV_REL_REC, ISV_VB_REC * 8);
  memcpy(out->rp, ho + ISV_REL_REC + ISV_VB_REC, ISV_RP_REC * 8);
  out->rank = ((const int32_t*)(ho + ISV_REL_REC + ISV_VB_REC + ISV_RP_REC))[1];
  out->status = ((const int32_t*)(ho + ISV_REL_REC + ISV_VB_REC + ISV_RP_REC + 1))[0];
  return ISV_OK;
}

// Mapped pinned block for the zero-copy event path (grow-only).
static isv_status ensure_mapped(isv_handle* h, size_t bytes) {
  if (h->mapped_bytes >= bytes) return ISV_OK;
  if (h->mapped) {
    cudaStreamSynchronize(h->stream);
    cudaFreeHost(h->mapped);
    h->mapped = nullptr;
    h->mapped_dev = nullptr;
    h->mapped_bytes = 0;
  }
  if (cudaHostAlloc((void**)&h->mapped, bytes + bytes / 2, cudaHostAllocMapped) != cudaSuccess) {
    cudaGetLastError();
    return ISV_ERR_ALLOC;
  }
  void* dp = nullptr;
  if (cudaHostGetDevicePointer(&dp, h->mapped, 0) != cudaSuccess) {
    cudaGetLastError();
    cudaFreeHost(h->mapped);
    h->mapped = nullptr;
    return ISV_ERR_CUDA;
  }
  h->mapped_dev = (char*)dp;
  h->mapped_bytes = bytes + bytes / 2;
  return ISV_OK;
}

// One MARGIN_OLD event, blocking.  Three routes (ISV_EVENT_MODE = 0 / 1 / 2 selects one for A/B measurements):
//   0 (default)  zero-copy: the event is packed into a mapped pinned block, the fused kernel reads it over PCIe, writes the
//                recovered factors back into the same block and publishes a completion word the host spins on: no copy
//                engine, no stream synchronisation on the critical path;
//   1            the fused kernel on a device-side mirror: one H2D, one launch, one D2H, cudaStreamSynchronize;
//   2            the batch kernels (five launches over three streams) on the device-side mirror -- the only route when
//                ProjectionFactor::sqrt_info is not a multiple of the identity.
isv_status isv_marg_event(isv_handle* h, const isv_fwd_in* fin, const isv_bwd_in* bin, isv_fwd_out* fout, isv_bwd_out* bout) {
  if (!h || !fin || !bin || !fout || !bout || fin->n_landmarks < 0) return ISV_ERR_BAD_ARG;
  if (!fin->pose0 || !fin->pose1 || !fin->ex_pose || !fin->prior_se3 || !fin->prior_rel) return ISV_ERR_BAD_ARG;
  if (!bin->pose_i || !bin->sb_i || !bin->pose_j || !bin->sb_j || !bin->prior_vb || !bin->preint) return ISV_ERR_BAD_ARG;
  const size_t L = (size_t)fin->n_landmarks;
  if (L > 0 && (!fin->inv_dep || !fin->pts_i || !fin->pts_j)) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  // in  = [lm_offset 2 (as int64)] [pose_fwd 14] [ex 7] [se3 48] [rel 48] [rp 5] [pose_bwd 14] [sb_bwd 18] [vb 90] [preint 467] [pad 1]
  //       [landmark components, L doubles each]
  // out = [se3 48] [pg 89] [rel 48] [vb 90] [rp 13] [rank 2 x i32 = 1] [status i32 = 1] [flag i32 = 1]
  // The fused routes (0, 1) pack the components the kernel reads next to each other -- x, y, [z,] inverse depth: 3 or 4 L
  // doubles -- so that the whole event is one contiguous block the kernel stages with a single bulk copy; the batch-kernel
  // route (2) keeps the ABI's six-component layout (x y z . . inverse depth).
  isv_batch_in probe;
  memset(&probe, 0, sizeof(probe));
  probe.n_windows = 1;
  probe.preint = bin->preint;
  const int mode = fused_eligible(h, &probe, ISV_RUN_BOTH) ? h->event_mode : 2;
  bool z_one = true;
  for (size_t k = 0; k < L; ++k) z_one &= (fin->pts_i[3 * k + 2] == 1.0);   // always, in the reference (src/System.cpp:346)
  const size_t n_rec = 14 + 7 + ISV_SE3_REC + ISV_REL_REC + ISV_RP_IN_REC + 14 + 18 + ISV_VB_REC + ISV_PREINT_REC + 1;   // 712
  const int lam_comp = mode == 2 ? 5 : (z_one ? 2 : 3);
  const size_t n_obs = (size_t)(lam_comp + 1) * L;
  const size_t n_in = (2 + n_rec + n_obs + 1) & ~(size_t)1;
  const size_t n_out = ISV_SE3_REC + ISV_PG_REC + ISV_REL_REC + ISV_VB_REC + ISV_RP_REC + 3;
  isv_status st;
  double* hp;     // host view of the block
  double* d;      // device view
  if (mode == 0) {
    st = ensure_mapped(h, (n_in + n_out) * sizeof(double));
    if (st != ISV_OK) return st;
    hp = (double*)h->mapped;
    d = (double*)h->mapped_dev;
  } else {
    st = ensure_pinned(h, (n_in + n_out) * sizeof(double));
    if (st != ISV_OK) return st;
    st = ensure_dbuf(h, (n_in + n_out + kScratchPerWindow + 64) * sizeof(double));
    if (st != ISV_OK) return st;
    hp = (double*)h->pinned;
    d = (double*)h->dbuf;
  }
  int64_t* off = (int64_t*)hp;
  off[0] = 0; off[1] = (int64_t)L;
  double* q = hp + 2;
  memcpy(q, fin->pose0, 56); memcpy(q + 7, fin->pose1, 56); memcpy(q + 14, fin->ex_pose, 56);
  memcpy(q + 21, fin->prior_se3, ISV_SE3_REC * 8); memcpy(q + 21 + ISV_SE3_REC, fin->prior_rel, ISV_REL_REC * 8);
  double* rp = q + 21 + ISV_SE3_REC + ISV_REL_REC;
  if (fin->prior_rp) memcpy(rp, fin->prior_rp, ISV_RP_IN_REC * 8); else memset(rp, 0, ISV_RP_IN_REC * 8);
  double* hb = rp + ISV_RP_IN_REC;
  memcpy(hb, bin->pose_i, 56); memcpy(hb + 7, bin->pose_j, 56);
  memcpy(hb + 14, bin->sb_i, 72); memcpy(hb + 23, bin->sb_j, 72);
  memcpy(hb + 32, bin->prior_vb, ISV_VB_REC * 8);
  memcpy(hb + 32 + ISV_VB_REC, bin->preint, ISV_PREINT_REC * 8);
  hb[32 + ISV_VB_REC + ISV_PREINT_REC] = 0.0;
  // the kernels read x_i, y_i, inv_dep and -- unless it is 1 everywhere -- z_i; pts_j is never read, so it is not packed
  double* obs = hp + 2 + n_rec;
  for (size_t k = 0; k < L; ++k) {
    obs[k] = fin->pts_i[3 * k];
    obs[L + k] = fin->pts_i[3 * k + 1];
    if (lam_comp != 2) obs[2 * L + k] = fin->pts_i[3 * k + 2];
    obs[(size_t)lam_comp * L + k] = fin->inv_dep[k];
  }
  cudaStream_t s = h->stream;
  double* dq = d + 2;
  double* db = dq + 21 + ISV_SE3_REC + ISV_REL_REC + ISV_RP_IN_REC;
  double* dout = d + n_in;
  isv_batch_in bi;
  memset(&bi, 0, sizeof(bi));
  bi.n_windows = 1;
  bi.ex_pose_shared = 1;
  bi.lm_offset = (const int64_t*)d;
  bi.lm_obs = d + 2 + n_rec;
  bi.lm_stride = (int64_t)L;
  bi.pose_fwd = dq;
  bi.ex_pose = dq + 14;
  bi.prior_se3 = dq + 21;
  bi.prior_rel = dq + 21 + ISV_SE3_REC;
  bi.prior_rp = dq + 21 + ISV_SE3_REC + ISV_REL_REC;
  bi.pose_bwd = db;
  bi.sb_bwd = db + 14;
  bi.prior_vb = db + 32;
  bi.preint = db + 32 + ISV_VB_REC;
  bi.flags = z_one ? ISV_IN_PTS_I_Z_ONE : 0;
  // what the fused kernel stages into shared memory with one bulk copy: the whole event if it fits, else the records
  static const bool no_stage = getenv("ISV_EVENT_NO_STAGE") != nullptr;   // A/B switch for measurements
  int stage_doubles = (int)((2 + n_rec + n_obs + 1) & ~(size_t)1);
  if (stage_doubles > kEvStageMaxDoubles) stage_doubles = (int)(2 + n_rec);
  if (no_stage) stage_doubles = 0;
  isv_batch_out bo;
  memset(&bo, 0, sizeof(bo));
  bo.se3_out = dout;
  bo.pg_out = bo.se3_out + ISV_SE3_REC;
  bo.rel_out = bo.pg_out + ISV_PG_REC;
  bo.vb_out = bo.rel_out + ISV_REL_REC;
  bo.rp_out = bo.vb_out + ISV_VB_REC;
  bo.rank = (int32_t*)(bo.rp_out + ISV_RP_REC);
  bo.status = (int32_t*)(bo.rp_out + ISV_RP_REC + 1);
  int32_t* dflag = (int32_t*)(bo.rp_out + ISV_RP_REC + 2);
  double* ho = hp + n_in;
  if (mode == 0) {
    volatile int32_t* hflag = (volatile int32_t*)(ho + (n_out - 1));
    const int32_t seq = ++h->event_seq;
    *hflag = seq - 1;
    __sync_synchronize();   // the packed event is in memory before the launch is submitted
    st = launch_fused(h, &bi, &bo, s, dflag, seq, nullptr, stage_doubles, lam_comp);
    if (st != ISV_OK) return st;
    // spin on the completion word; every ~4 k polls make sure the stream has not died under us
    for (unsigned spins = 1;; ++spins) {
      if (*hflag == seq) break;
      if ((spins & 0xfff) == 0) {
        const cudaError_t e = cudaStreamQuery(s);
        if (e == cudaSuccess) {
          if (*hflag == seq) break;
          return ISV_ERR_CUDA;   // the kernel ended without publishing
        }
        if (e != cudaErrorNotReady) { cudaGetLastError(); return ISV_ERR_CUDA; }
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    __sync_synchronize();
  } else {
    ISV_CUDA(cudaMemcpyAsync(d, hp, n_in * sizeof(double), cudaMemcpyHostToDevice, s));
    if (mode == 1) {
      st = launch_fused(h, &bi, &bo, s, nullptr, 0, nullptr, stage_doubles, lam_comp);
    } else {
      ISV_CUDA(cudaMemsetAsync(bo.rank, 0, 8, s));
      st = launch_batch(h, &bi, &bo, ISV_RUN_BOTH, s, dout + n_out);
    }
    if (st != ISV_OK) return st;
    ISV_CUDA(cudaMemcpyAsync(ho, dout, n_out * sizeof(double), cudaMemcpyDeviceToHost, s));
    ISV_CUDA(cudaStreamSynchronize(s));
  }
  const double* o = ho;
  memcpy(fout->se3, o, ISV_SE3_REC * 8); o += ISV_SE3_REC;
  memcpy(fout->pg, o, ISV_PG_REC * 8); o += ISV_PG_REC;
  memcpy(bout->rel, o, ISV_REL_REC * 8); o += ISV_REL_REC;
  memcpy(bout->vb, o, ISV_VB_REC * 8); o += ISV_VB_REC;
  memcpy(bout->rp, o, ISV_RP_REC * 8); o += ISV_RP_REC;
  fout->rank = ((const int32_t*)o)[0];
  bout->rank = ((const int32_t*)o)[1];
  fout->status = bout->status = ((const int32_t*)(o + 1))[0];
  return ISV_OK;
}

isv_status isv_test_fused_stamps(isv_handle* h, const isv_batch_in* in, const isv_batch_out* out, int64_t* stamps_dev) {
  if (!h || !stamps_dev) return ISV_ERR_BAD_ARG;
  isv_status st = check_batch(in, out, ISV_RUN_BOTH);
  if (st != ISV_OK) return st;
  if (!in->preint || in->n_windows < 1) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  return launch_fused(h, in, out, h->stream, nullptr, 0, (long long*)stamps_dev);
}

isv_status isv_test_event_latency(isv_handle* h, const isv_fwd_in* fin, const isv_bwd_in* bin, isv_fwd_out* fout, isv_bwd_out* bout,
                                  int iters, double* us) {
  if (!h || iters < 0 || (iters > 0 && !us)) return ISV_ERR_BAD_ARG;
  for (int i = 0; i < iters; ++i) {
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const isv_status st = isv_marg_event(h, fin, bin, fout, bout);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (st != ISV_OK) return st;
    us[i] = (double)(t1.tv_sec - t0.tv_sec) * 1e6 + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-3;
  }
  return ISV_OK;
}

isv_status isv_pose_plus_batch(isv_handle* h, int64_t n, const double* x, const double* delta, double* x_plus_delta) {
  if (!h || n < 0 || (n > 0 && (!x || !delta || !x_plus_delta))) return ISV_ERR_BAD_ARG;
  if (n == 0) return ISV_OK;
  ISV_CUDA(cudaSetDevice(h->device));
  const long long grid = (n + 127) / 128;
  if (grid > 0x7fffffffLL) return ISV_ERR_BAD_ARG;
  pose_plus_kernel<<<(unsigned)grid, 128, 0, h->stream>>>((long long)n, x, delta, x_plus_delta);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

void isv_pose_plus_jacobian(double* jacobian) {
  if (!jacobian) return;
  for (int i = 0; i < 42; ++i) jacobian[i] = 0.0;
  for (int i = 0; i < 6; ++i) jacobian[6 * i + i] = 1.0;   // 7 x 6 row-major: top 6 rows identity, last row zero
}

}  // extern "C"

// ---- unit-test hook for the warp linear algebra (tests/test_linalg_gpu.py) ---------------------
namespace isv {
__global__ void psd_eig_test_kernel(int nb, int n, const double* A, double* G, double* lam, int* info) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= nb) return;
  double* As = smem + warp * (2 * n * n + 2 * n);
  double* Gs = As + n * n;
  double* d = Gs + n * n;
  for (int i = lane; i < n * n; i += 32) { As[i] = A[(size_t)b * n * n + i]; Gs[i] = 0.0; }
  __syncwarp();
  int r = w_pivoted_cholesky_rows(As, n, n, Gs, n, d, lane);
  int sweeps = (n <= 8) ? w_onesided_jacobi_rows<4, 2>(Gs, n, r, n, d, lane)
               : (n <= 24) ? w_onesided_jacobi_rows<4, 6>(Gs, n, r, n, d, lane)
                           : w_onesided_jacobi_rows<2, 32>(Gs, n, r, n, d, lane);
  if (n == 21) {}
  for (int i = lane; i < n * n; i += 32) G[(size_t)b * n * n + i] = Gs[i];
  for (int i = lane; i < n; i += 32) lam[(size_t)b * n + i] = i < r ? d[i] : 0.0;
  if (lane == 0) { info[2 * b] = r; info[2 * b + 1] = sweeps; }
}
}  // namespace isv

// ---- unit-test hook: the lean reciprocal square root / reciprocal of isv_device_math.cuh ------------------------------
namespace isv {
__global__ void fast_special_kernel(int n, const double* __restrict__ x, double* __restrict__ rs, double* __restrict__ rc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { rs[i] = fast_rsqrt(x[i]); rc[i] = fast_rcp(x[i]); }
}
}  // namespace isv
extern "C" isv_status isv_test_fast_special(isv_handle* h, int n, const double* x, double* rsqrt_out, double* rcp_out) {
  if (!h || n < 0 || (n > 0 && (!x || !rsqrt_out || !rcp_out))) return ISV_ERR_BAD_ARG;
  if (n == 0) return ISV_OK;
  ISV_CUDA(cudaSetDevice(h->device));
  isv_status st = ensure_dbuf(h, (size_t)3 * n * sizeof(double));
  if (st != ISV_OK) return st;
  double* d = (double*)h->dbuf;
  ISV_CUDA(cudaMemcpyAsync(d, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  fast_special_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, d, d + n, d + 2 * (size_t)n);
  ISV_CUDA(cudaMemcpyAsync(rsqrt_out, d + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  ISV_CUDA(cudaMemcpyAsync(rcp_out, d + 2 * (size_t)n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  ISV_CUDA(cudaStreamSynchronize(h->stream));
  return ISV_OK;
}

extern "C" isv_status isv_test_psd_eig(isv_handle* h, int nb, int n, const double* A, double* G, double* lam,
                                       int32_t* info) {
  if (!h || nb <= 0 || n < 1 || n > 63 || !A || !G || !lam || !info) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  double *dA, *dG, *dl;
  int* di;
  size_t mb = sizeof(double) * (size_t)nb * n * n;
  ISV_CUDA(cudaMalloc(&dA, mb));
  ISV_CUDA(cudaMalloc(&dG, mb));
  ISV_CUDA(cudaMalloc(&dl, sizeof(double) * (size_t)nb * n));
  ISV_CUDA(cudaMalloc(&di, sizeof(int) * 2 * (size_t)nb));
  ISV_CUDA(cudaMemcpyAsync(dA, A, mb, cudaMemcpyHostToDevice, h->stream));
  const int wpc = 2;
  size_t sm = wpc * (2 * n * n + 2 * n) * sizeof(double);
  cudaFuncSetAttribute(psd_eig_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  psd_eig_test_kernel<<<(nb + wpc - 1) / wpc, 32 * wpc, sm, h->stream>>>(nb, n, dA, dG, dl, di);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  ISV_CUDA(cudaMemcpyAsync(G, dG, mb, cudaMemcpyDeviceToHost, h->stream));
  ISV_CUDA(cudaMemcpyAsync(lam, dl, sizeof(double) * (size_t)nb * n, cudaMemcpyDeviceToHost, h->stream));
  ISV_CUDA(cudaMemcpyAsync(info, di, sizeof(int) * 2 * (size_t)nb, cudaMemcpyDeviceToHost, h->stream));
  ISV_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(dA); cudaFree(dG); cudaFree(dl); cudaFree(di);
  return ISV_OK;
}

// ---- initFactorGraph sparsification tail ------------------------------------------------------------
extern "C" isv_status isv_init_sparsify_batch(isv_handle* h, const isv_init_in* in, const isv_init_out* out) {
  if (!h || !in || !out || in->n_windows < 0) return ISV_ERR_BAD_ARG;
  if (!in->poses || !in->sbs || !in->preint || !out->rel_out || !out->se3_out || !out->vb_out || !out->rank)
    return ISV_ERR_BAD_ARG;
  const int V = h->cfg.vo_size;
  if (V < 2 || V > 10) return ISV_ERR_BAD_ARG;
  if (in->n_windows == 0) return ISV_OK;
  ISV_CUDA(cudaSetDevice(h->device));
  const size_t sm = init_smem_doubles(V) * sizeof(double);
  ISV_CUDA(cudaFuncSetAttribute(init_sparsify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  init_sparsify_kernel<<<in->n_windows, 32, sm, h->stream>>>(in->n_windows, V, in->poses, in->sbs, in->preint,
                                                           out->rel_out, out->se3_out, out->vb_out, out->rank,
                                                           out->status, h->dcfg);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

extern "C" isv_status isv_init_sparsify_host(isv_handle* h, const isv_init_in* in, const isv_init_out* out) {
  if (!h || !in || !out || in->n_windows < 0) return ISV_ERR_BAD_ARG;
  if (!in->poses || !in->sbs || !in->preint || !out->rel_out || !out->se3_out || !out->vb_out || !out->rank)
    return ISV_ERR_BAD_ARG;
  const size_t n = (size_t)in->n_windows, V = (size_t)h->cfg.vo_size, D = sizeof(double);
  if (V < 2 || V > 10) return ISV_ERR_BAD_ARG;
  if (n == 0) return ISV_OK;
  ISV_CUDA(cudaSetDevice(h->device));
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  const size_t o_p = carve(n * V * 7 * D), o_s = carve(n * V * 9 * D), o_pre = carve(n * (V - 1) * ISV_PREINT_REC * D);
  const size_t o_rel = carve(n * (V - 1) * ISV_REL_REC * D), o_se3 = carve(n * ISV_SE3_REC * D);
  const size_t o_vb = carve(n * ISV_VB_REC * D), o_rank = carve(n * 4), o_st = carve(n * 4);
  isv_status st = ensure_dbuf(h, off);
  if (st != ISV_OK) return st;
  char* d = h->dbuf;
  cudaStream_t s = h->stream;
  ISV_CUDA(cudaMemcpyAsync(d + o_p, in->poses, n * V * 7 * D, cudaMemcpyHostToDevice, s));
  ISV_CUDA(cudaMemcpyAsync(d + o_s, in->sbs, n * V * 9 * D, cudaMemcpyHostToDevice, s));
  ISV_CUDA(cudaMemcpyAsync(d + o_pre, in->preint, n * (V - 1) * ISV_PREINT_REC * D, cudaMemcpyHostToDevice, s));
  isv_init_in di = {in->n_windows, (const double*)(d + o_p), (const double*)(d + o_s), (const double*)(d + o_pre)};
  isv_init_out dout = {(double*)(d + o_rel), (double*)(d + o_se3), (double*)(d + o_vb), (int32_t*)(d + o_rank),
                       (int32_t*)(d + o_st)};
  st = isv_init_sparsify_batch(h, &di, &dout);
  if (st != ISV_OK) return st;
  ISV_CUDA(cudaMemcpyAsync(out->rel_out, dout.rel_out, n * (V - 1) * ISV_REL_REC * D, cudaMemcpyDeviceToHost, s));
  ISV_CUDA(cudaMemcpyAsync(out->se3_out, dout.se3_out, n * ISV_SE3_REC * D, cudaMemcpyDeviceToHost, s));
  ISV_CUDA(cudaMemcpyAsync(out->vb_out, dout.vb_out, n * ISV_VB_REC * D, cudaMemcpyDeviceToHost, s));
  ISV_CUDA(cudaMemcpyAsync(out->rank, dout.rank, n * 4, cudaMemcpyDeviceToHost, s));
  if (out->status) ISV_CUDA(cudaMemcpyAsync(out->status, dout.status, n * 4, cudaMemcpyDeviceToHost, s));
  ISV_CUDA(cudaStreamSynchronize(s));
  return ISV_OK;
}

// ---- IMU pre-integration ------------------------------------------------------------------------------
extern "C" isv_status isv_preintegrate_batch(isv_handle* h, const isv_preint_in* in, double* preint_out) {
  if (!h || !in || !preint_out || in->n < 0 || in->k_max < 0 || !in->imu_init) return ISV_ERR_BAD_ARG;
  if (in->k_max > 0 && !in->imu_raw) return ISV_ERR_BAD_ARG;
  if (in->n == 0) return ISV_OK;
  ISV_CUDA(cudaSetDevice(h->device));
  const size_t sm = kWarpsPerCta * kPreSmemPerWarp * sizeof(double);
  ISV_CUDA(cudaFuncSetAttribute(preintegrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  NoiseCfg nz{h->cfg.acc_n, h->cfg.gyr_n, h->cfg.acc_w, h->cfg.gyr_w};
  preintegrate_kernel<<<(in->n + kWarpsPerCta - 1) / kWarpsPerCta, kThreads, sm, h->stream>>>(
      in->n, in->k_max, in->k_count, in->imu_raw, in->imu_init, preint_out, nz);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

extern "C" isv_status isv_preintegrate_host(isv_handle* h, const isv_preint_in* in, double* preint_out) {
  if (!h || !in || !preint_out || in->n < 0 || in->k_max < 0 || !in->imu_init) return ISV_ERR_BAD_ARG;
  if (in->k_max > 0 && !in->imu_raw) return ISV_ERR_BAD_ARG;
  const size_t n = (size_t)in->n, K = (size_t)in->k_max, D = sizeof(double);
  if (n == 0) return ISV_OK;
  if (in->k_count)   // host pointers here: a count outside [0, k_max] is a caller bug, not something to clamp silently
    for (size_t w = 0; w < n; ++w)
      if (in->k_count[w] < 0 || in->k_count[w] > in->k_max) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  const size_t o_raw = carve(n * K * 7 * D), o_init = carve(n * 12 * D), o_k = carve(in->k_count ? n * 4 : 0);
  const size_t o_out = carve(n * ISV_PREINT_REC * D);
  isv_status st = ensure_dbuf(h, off);
  if (st != ISV_OK) return st;
  char* d = h->dbuf;
  cudaStream_t s = h->stream;
  if (K) ISV_CUDA(cudaMemcpyAsync(d + o_raw, in->imu_raw, n * K * 7 * D, cudaMemcpyHostToDevice, s));
  ISV_CUDA(cudaMemcpyAsync(d + o_init, in->imu_init, n * 12 * D, cudaMemcpyHostToDevice, s));
  if (in->k_count) ISV_CUDA(cudaMemcpyAsync(d + o_k, in->k_count, n * 4, cudaMemcpyHostToDevice, s));
  isv_preint_in di = {in->n, in->k_max, in->k_count ? (const int32_t*)(d + o_k) : nullptr, (const double*)(d + o_raw),
                      (const double*)(d + o_init)};
  st = isv_preintegrate_batch(h, &di, (double*)(d + o_out));
  if (st != ISV_OK) return st;
  ISV_CUDA(cudaMemcpyAsync(preint_out, d + o_out, n * ISV_PREINT_REC * D, cudaMemcpyDeviceToHost, s));
  ISV_CUDA(cudaStreamSynchronize(s));
  return ISV_OK;
}

// ---- ceres Evaluate contract, batched -----------------------------------------------------------------
static bool pb_ok(const isv_param_blocks* pb) {
  if (!pb || pb->n_pose < 0 || pb->n_speed_bias < 0 || pb->n_ex_pose < 0 || pb->n_feature < 0) return false;
  if ((pb->n_pose && !pb->pose) || (pb->n_speed_bias && !pb->speed_bias) || (pb->n_ex_pose && !pb->ex_pose) ||
      (pb->n_feature && !pb->feature))
    return false;
  return true;
}

static isv_status eval_projection_on(isv_handle* h, cudaStream_t stream, const isv_param_blocks* pb,
                                     const isv_proj_factors* f, const isv_proj_eval* out, int32_t* status) {
  if (!h || !pb_ok(pb) || !f || !out || f->n < 0 || f->stride < f->n) return ISV_ERR_BAD_ARG;
  if (f->n == 0) return ISV_OK;
  if (!f->idx || !f->obs || !out->residuals || !(f->cauchy_a >= 0.0)) return ISV_ERR_BAD_ARG;
  const long long per_cta = 32LL * kEvalWarps;
  const long long grid = (f->n + per_cta - 1) / per_cta;
  if (grid > 0x7fffffffLL) return ISV_ERR_BAD_ARG;
  if (f->td_obs) {
    if (!f->td || f->n_td < 1) return ISV_ERR_BAD_ARG;
    eval_projection_kernel<true><<<(unsigned)grid, kEvalThreads, 0, stream>>>(*pb, *f, *out, h->dcfg, status);
  } else {
    eval_projection_kernel<false><<<(unsigned)grid, kEvalThreads, 0, stream>>>(*pb, *f, *out, h->dcfg, status);
  }
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

static isv_status eval_imu_on(isv_handle* h, cudaStream_t stream, const isv_param_blocks* pb, const isv_imu_factors* f,
                              const isv_imu_eval* out, int32_t* status) {
  if (!h || !pb_ok(pb) || !f || !out || f->n < 0) return ISV_ERR_BAD_ARG;
  if (f->n == 0) return ISV_OK;
  if (!f->idx || !f->preint || !out->residuals) return ISV_ERR_BAD_ARG;
  const size_t sm = kEvalWarps * kImuEvalSmem * sizeof(double);
  ISV_CUDA(cudaFuncSetAttribute(eval_imu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  eval_imu_kernel<<<(f->n + kEvalWarps - 1) / kEvalWarps, kEvalThreads, sm, stream>>>(*pb, *f, *out, h->dcfg, status);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

static isv_status eval_small_on(isv_handle* h, cudaStream_t stream, const isv_param_blocks* pb,
                                const isv_small_factors* f, const isv_small_eval* out, int32_t* status) {
  if (!h || !pb_ok(pb) || !f || !out) return ISV_ERR_BAD_ARG;
  if (f->n_rel < 0 || f->n_se3 < 0 || f->n_vb < 0 || f->n_rp < 0 || f->n_yaw < 0 || !(f->cauchy_a >= 0.0))
    return ISV_ERR_BAD_ARG;
  if ((f->n_rel && (!f->rel_idx || !f->rel_rec || !out->rel_res)) || (f->n_se3 && (!f->se3_idx || !f->se3_rec || !out->se3_res)) ||
      (f->n_vb && (!f->vb_idx || !f->vb_rec || !out->vb_res)) || (f->n_rp && (!f->rp_idx || !f->rp_rec || !out->rp_res)) ||
      (f->n_yaw && (!f->yaw_idx || !f->yaw_rec || !out->yaw_res)))
    return ISV_ERR_BAD_ARG;
  const long long total = (long long)f->n_rel + f->n_se3 + f->n_vb + f->n_rp + f->n_yaw;
  if (total == 0) return ISV_OK;
  eval_small_kernel<<<(unsigned)((total + kEvalThreads - 1) / kEvalThreads), kEvalThreads, 0, stream>>>(*pb, *f, *out,
                                                                                                       status);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

extern "C" isv_status isv_eval_projection_batch(isv_handle* h, const isv_param_blocks* pb, const isv_proj_factors* f,
                                                const isv_proj_eval* out, int32_t* status) {
  if (!h) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  return eval_projection_on(h, h->stream, pb, f, out, status);
}

extern "C" isv_status isv_eval_imu_batch(isv_handle* h, const isv_param_blocks* pb, const isv_imu_factors* f,
                                         const isv_imu_eval* out, int32_t* status) {
  if (!h) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  return eval_imu_on(h, h->stream, pb, f, out, status);
}

extern "C" isv_status isv_eval_small_batch(isv_handle* h, const isv_param_blocks* pb, const isv_small_factors* f,
                                           const isv_small_eval* out, int32_t* status) {
  if (!h) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  return eval_small_on(h, h->stream, pb, f, out, status);
}

// All factor classes of one problemSolve() iteration.  The IMU and prior-factor kernels are
// latency-bound (a few thousand short serial chains): they are forked onto two side streams and
// overlap the HBM-bound projection kernel, then joined back into the handle's stream.
extern "C" isv_status isv_eval_problem(isv_handle* h, const isv_param_blocks* pb, const isv_proj_factors* pf,
                                       const isv_proj_eval* po, const isv_imu_factors* mf, const isv_imu_eval* mo,
                                       const isv_small_factors* sf, const isv_small_eval* so, int32_t* status) {
  if (!h || (pf && !po) || (mf && !mo) || (sf && !so)) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  const bool fork = (mf && mf->n > 0) || (sf != nullptr);
  if (fork) {
    ISV_CUDA(cudaEventRecord(h->aux_ev[0], h->stream));
    ISV_CUDA(cudaStreamWaitEvent(h->aux[0], h->aux_ev[0], 0));
    ISV_CUDA(cudaStreamWaitEvent(h->aux[1], h->aux_ev[0], 0));
  }
  isv_status st = ISV_OK;
  if (mf) st = eval_imu_on(h, h->aux[0], pb, mf, mo, status);
  if (st == ISV_OK && sf) st = eval_small_on(h, h->aux[1], pb, sf, so, status);
  if (st == ISV_OK && pf) st = eval_projection_on(h, h->stream, pb, pf, po, status);
  if (fork) {   // always join, also on error, so the streams stay ordered
    cudaEventRecord(h->aux_ev[1], h->aux[0]);
    cudaEventRecord(h->aux_ev[2], h->aux[1]);
    cudaStreamWaitEvent(h->stream, h->aux_ev[1], 0);
    cudaStreamWaitEvent(h->stream, h->aux_ev[2], 0);
  }
  return st;
}

// ---- device-resident sequence state ---------------------------------------------------------------------
struct isv_seq {
  SeqView v;
  char* slab;
  size_t bytes;
  double* rel_init;   // [n][V-1][48] scratch of isv_seq_init
  double* gram;       // [n][42]
};

extern "C" isv_status isv_seq_create(isv_handle* h, int n, isv_seq** out) {
  if (!h || !out || n <= 0) return ISV_ERR_BAD_ARG;
  *out = nullptr;
  const int V = h->cfg.vo_size;
  if (V < 2 || V > 10) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  isv_seq* s = new (std::nothrow) isv_seq();
  if (!s) return ISV_ERR_ALLOC;
  memset(s, 0, sizeof(*s));
  const size_t D = sizeof(double), N = (size_t)n;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  const size_t o_rel = carve(V * N * ISV_REL_REC * D), o_se3 = carve(N * ISV_SE3_REC * D), o_vb = carve(N * ISV_VB_REC * D);
  const size_t o_rp = carve(V * N * ISV_RP_REC * D), o_rpv = carve(V * N * 4), o_rpin = carve(N * ISV_RP_IN_REC * D);
  const size_t o_acc = carve(N * ISV_ACC_REC * D), o_cnt = carve(N * 4);
  const size_t o_se3o = carve(N * ISV_SE3_REC * D), o_pgo = carve(N * ISV_PG_REC * D), o_relo = carve(N * ISV_REL_REC * D);
  const size_t o_vbo = carve(N * ISV_VB_REC * D), o_rpo = carve(N * ISV_RP_REC * D), o_rank = carve(N * 8), o_st = carve(N * 4);
  const size_t o_ri = carve(N * (V - 1) * ISV_REL_REC * D), o_gram = carve(N * kScratchPerWindow * D);
  if (cudaMalloc(&s->slab, off) != cudaSuccess) {
    cudaGetLastError();
    delete s;
    return ISV_ERR_ALLOC;
  }
  s->bytes = off;
  cudaMemsetAsync(s->slab, 0, off, h->stream);
  char* d = s->slab;
  s->v = SeqView{n, V, (double*)(d + o_rel), (double*)(d + o_se3), (double*)(d + o_vb), (double*)(d + o_rp),
                 (int32_t*)(d + o_rpv), (double*)(d + o_rpin), (double*)(d + o_acc), (int32_t*)(d + o_cnt),
                 (double*)(d + o_se3o), (double*)(d + o_pgo), (double*)(d + o_relo), (double*)(d + o_vbo),
                 (double*)(d + o_rpo), (int32_t*)(d + o_rank), (int32_t*)(d + o_st)};
  s->rel_init = (double*)(d + o_ri);
  s->gram = (double*)(d + o_gram);
  *out = s;
  return ISV_OK;
}

extern "C" void isv_seq_destroy(isv_handle* h, isv_seq* s) {
  if (!s) return;
  if (h) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
  }
  if (s->slab) cudaFree(s->slab);
  delete s;
}

extern "C" isv_status isv_seq_init(isv_handle* h, isv_seq* s, const isv_init_in* in, int32_t* rank) {
  if (!h || !s || !in || in->n_windows != s->v.n) return ISV_ERR_BAD_ARG;
  isv_init_out o = {s->rel_init, s->v.se3, s->v.vb, rank ? rank : s->v.rank, s->v.status};
  ISV_CUDA(cudaMemsetAsync(s->v.status, 0, sizeof(int32_t) * (size_t)s->v.n, h->stream));
  isv_status st = isv_init_sparsify_batch(h, in, &o);
  if (st != ISV_OK) return st;
  seq_install_kernel<<<s->v.n, 64, 0, h->stream>>>(s->v, s->rel_init);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

extern "C" isv_status isv_seq_update(isv_handle* h, isv_seq* s, const isv_seq_update_in* in) {
  if (!h || !s || !in || !in->old_P || !in->old_R || !in->old_vb || !in->pose || !in->speed_bias) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  const long long total = (long long)s->v.n * (2 + (s->v.V - 1) + s->v.V);
  seq_update_kernel<<<(unsigned)((total + 127) / 128), 128, 0, h->stream>>>(s->v, in->old_P, in->old_R, in->old_vb, in->pose,
                                                                          in->speed_bias);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

extern "C" isv_status isv_seq_yaw(isv_handle* h, isv_seq* s, const double* old_R0, const double* pose0,
                                  double* rot_diff_out) {
  if (!h || !s || !old_R0 || !pose0) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  seq_yaw_kernel<<<(s->v.n + 127) / 128, 128, 0, h->stream>>>(s->v, old_R0, pose0, rot_diff_out);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

extern "C" isv_status isv_seq_marginalize(isv_handle* h, isv_seq* s, const isv_seq_frame* f, double pg_cut_distance,
                                          double* kf_out, int32_t* kf_flag) {
  if (!h || !s || !f) return ISV_ERR_BAD_ARG;
  const bool pg = f->ts || f->Ri || f->ti;
  if (pg && (!f->ts || !f->Ri || !f->ti)) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  const SeqView& v = s->v;
  isv_batch_in bi;
  memset(&bi, 0, sizeof(bi));
  bi.n_windows = v.n;
  bi.ex_pose_shared = f->ex_pose_shared;
  bi.lm_offset = f->lm_offset;
  bi.lm_obs = f->lm_obs;
  bi.lm_stride = f->lm_stride;
  bi.pose_fwd = f->pose_fwd;
  bi.ex_pose = f->ex_pose;
  bi.prior_se3 = v.se3;                                      // vioPosePriorEdge
  bi.prior_rel = v.rel + (size_t)v.n * ISV_REL_REC;          // vioRelativePoseEdges[1]
  bi.prior_rp = v.rp_in;                                     // vioRollPitchEdges[0] if its index is 0
  bi.pose_bwd = f->pose_bwd;
  bi.sb_bwd = f->sb_bwd;
  bi.prior_vb = v.vb;                                        // vioVBPrior
  bi.preint = f->preint;
  isv_batch_out bo = {v.se3_out, v.pg_out, v.rel_out, v.vb_out, v.rp_out, v.rank, v.status};
  isv_status st = check_batch(&bi, &bo, ISV_RUN_BOTH);
  if (st != ISV_OK) return st;
  st = launch_batch(h, &bi, &bo, ISV_RUN_BOTH, h->stream, s->gram);
  if (st != ISV_OK) return st;
  if (pg) {
    const int wpc = 4;
    seq_pg_kernel<<<(v.n + wpc - 1) / wpc, 32 * wpc, wpc * kPgSmemPerWarp * sizeof(double), h->stream>>>(
        v, f->ts, f->Ri, f->ti, pg_cut_distance, kf_out, kf_flag);
    ++h->launches;
  }
  seq_rotate_kernel<<<v.n, 64, 0, h->stream>>>(v);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

static isv_status seq_copy(isv_handle* h, isv_seq* s, const isv_seq_host* x, bool to_host) {
  if (!h || !s || !x) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  const SeqView& v = s->v;
  const size_t D = sizeof(double), N = (size_t)v.n, V = (size_t)v.V;
  struct Item { void* host; void* dev; size_t bytes; };
  const Item items[] = {
      {x->rel, v.rel, V * N * ISV_REL_REC * D}, {x->se3, v.se3, N * ISV_SE3_REC * D}, {x->vb, v.vb, N * ISV_VB_REC * D},
      {x->rp, v.rp, V * N * ISV_RP_REC * D}, {x->rp_valid, v.rp_valid, V * N * 4}, {x->acc, v.acc, N * ISV_ACC_REC * D},
      {x->pg_count, v.pg_count, N * 4}, {x->last_se3, v.se3_out, N * ISV_SE3_REC * D}, {x->last_pg, v.pg_out, N * ISV_PG_REC * D},
      {x->last_rel, v.rel_out, N * ISV_REL_REC * D}, {x->last_vb, v.vb_out, N * ISV_VB_REC * D},
      {x->last_rp, v.rp_out, N * ISV_RP_REC * D}, {x->last_rank, v.rank, N * 8}, {x->last_status, v.status, N * 4}};
  for (const Item& it : items) {
    if (!it.host) continue;
    if (to_host) ISV_CUDA(cudaMemcpyAsync(it.host, it.dev, it.bytes, cudaMemcpyDeviceToHost, h->stream));
    else ISV_CUDA(cudaMemcpyAsync(it.dev, it.host, it.bytes, cudaMemcpyHostToDevice, h->stream));
  }
  if (!to_host && x->rp && x->rp_valid) {
    // rebuild the packed slot-0 record the forward kernel reads
    // (done on the host side of the copy: valid flag + sqrt_info of slot 0)
    double* tmp = (double*)malloc(N * ISV_RP_IN_REC * D);
    if (!tmp) return ISV_ERR_ALLOC;
    for (size_t q = 0; q < N; ++q) {
      const int ok = x->rp_valid[q] != 0;
      tmp[q * ISV_RP_IN_REC] = ok ? 1.0 : 0.0;
      for (int k = 0; k < 4; ++k) tmp[q * ISV_RP_IN_REC + 1 + k] = ok ? x->rp[q * ISV_RP_REC + 9 + k] : 0.0;
    }
    cudaError_t e = cudaMemcpyAsync(v.rp_in, tmp, N * ISV_RP_IN_REC * D, cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    free(tmp);
    if (e != cudaSuccess) return ISV_ERR_CUDA;
  }
  ISV_CUDA(cudaStreamSynchronize(h->stream));
  return ISV_OK;
}

extern "C" isv_status isv_seq_export_host(isv_handle* h, isv_seq* s, const isv_seq_host* out) {
  return seq_copy(h, s, out, true);
}
extern "C" isv_status isv_seq_import_host(isv_handle* h, isv_seq* s, const isv_seq_host* in) {
  return seq_copy(h, s, in, false);
}

// ---- generic marginalization (MarginalizationInfo engine) ---------------------------------------------------
static isv_status marg_generic_check(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out,
                                     bool full) {
  if (!h || !in || !out) return ISV_ERR_BAD_ARG;
  if (in->n_problems < 1 || in->pos < 1 || in->m_dense < 0 || in->m_diag < 0 || in->m_dense > kMgMaxDense) return ISV_ERR_BAD_ARG;
  if (in->m_dense + in->m_diag >= in->pos || in->n_factors < 0 || !(in->eps >= 0.0)) return ISV_ERR_BAD_ARG;
  if (in->n_factors > 0 && (!in->factors || !in->blocks || !in->values)) return ISV_ERR_BAD_ARG;
  if (!out->A || !out->b) return ISV_ERR_BAD_ARG;
  if (full && (!out->A_red || !out->b_red || !out->linearized_jacobians || !out->linearized_residuals || !out->rank))
    return ISV_ERR_BAD_ARG;
  return ISV_OK;
}

extern "C" isv_status isv_build_normal_equations(isv_handle* h, const isv_marg_generic_in* in,
                                                 const isv_marg_generic_out* out) {
  isv_status st = marg_generic_check(h, in, out, false);
  if (st != ISV_OK) return st;
  ISV_CUDA(cudaSetDevice(h->device));
  const size_t np = (size_t)in->n_problems, pos = (size_t)in->pos;
  ISV_CUDA(cudaMemsetAsync(out->A, 0, np * pos * pos * sizeof(double), h->stream));
  ISV_CUDA(cudaMemsetAsync(out->b, 0, np * pos * sizeof(double), h->stream));
  if (out->status) ISV_CUDA(cudaMemsetAsync(out->status, 0, np * sizeof(int32_t), h->stream));
  if (in->n_factors > 0) {
    const long long grid = (in->n_factors + kNeWarps - 1) / kNeWarps;
    if (grid > 0x7fffffffLL) return ISV_ERR_BAD_ARG;
    ne_build_kernel<<<(unsigned)grid, 32 * kNeWarps, kNeWarps * kNeSmemPerWarp * sizeof(double), h->stream>>>(
        *in, out->A, out->b, out->status);
    ++h->launches;
  }
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

static isv_status marginalize_generic_impl(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out,
                                           int schur_only);

extern "C" isv_status isv_marginalize_generic(isv_handle* h, const isv_marg_generic_in* in,
                                              const isv_marg_generic_out* out) {
  return marginalize_generic_impl(h, in, out, 0);
}

// normal equations + Schur complement only: with every feature in the diagonal block and m_dense = 0 this
// is the reduced camera system ceres' DENSE_SCHUR solves (src/estimator.cpp:1124), built on the GPU
extern "C" isv_status isv_reduced_system(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out) {
  if (!out || !out->A_red || !out->b_red || !out->rank) return ISV_ERR_BAD_ARG;
  isv_marg_generic_out o = *out;
  if (!o.linearized_jacobians) o.linearized_jacobians = o.A_red;   // not written in this mode
  if (!o.linearized_residuals) o.linearized_residuals = o.b_red;
  return marginalize_generic_impl(h, in, &o, 1);
}

static isv_status schur_eig_impl(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out,
                                 int schur_only);

static isv_status marginalize_generic_impl(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out,
                                           int schur_only) {
  isv_status st = marg_generic_check(h, in, out, true);
  if (st != ISV_OK) return st;
  st = isv_build_normal_equations(h, in, out);
  if (st != ISV_OK) return st;
  return schur_eig_impl(h, in, out, schur_only);
}

extern "C" isv_status isv_schur_eig(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out,
                                    int32_t schur_only) {
  if (schur_only) {
    if (!out || !out->A_red || !out->b_red || !out->rank) return ISV_ERR_BAD_ARG;
    isv_marg_generic_out o = *out;
    if (!o.linearized_jacobians) o.linearized_jacobians = o.A_red;
    if (!o.linearized_residuals) o.linearized_residuals = o.b_red;
    isv_status st = marg_generic_check(h, in, &o, true);
    return st != ISV_OK ? st : schur_eig_impl(h, in, &o, 1);
  }
  isv_status st = marg_generic_check(h, in, out, true);
  return st != ISV_OK ? st : schur_eig_impl(h, in, out, 0);
}

// ---- MarginalizationFactor (the previous prior as a residual block) ----------------------------------------
static bool prior_ok(const isv_marg_prior* p) {
  return p && p->n >= 1 && p->n_blocks >= 1 && p->blocks && p->linearized_jacobians && p->linearized_residuals && p->x0 &&
         p->x;
}

extern "C" isv_status isv_eval_marg_prior(isv_handle* h, const isv_marg_prior* prior, double* residuals, double* jacobians,
                                          int32_t* status) {
  if (!h || !prior_ok(prior) || !residuals) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  const size_t sm = (size_t)prior->n * sizeof(double) + ((size_t)prior->n_blocks + 2) * sizeof(int);
  if (sm > 200 * 1024) return ISV_ERR_BAD_ARG;
  if (sm > 48 * 1024)
    ISV_CUDA(cudaFuncSetAttribute(marg_prior_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  // rows of the residual split over the CTAs; with Jacobians requested, enough CTAs to stream the re-layout
  int grid = (prior->n + kPriorThreads - 1) / kPriorThreads;
  if (jacobians) grid = std::max(grid, std::min(2 * 148, (prior->n * prior->n) / (4 * kPriorThreads) + 1));
  marg_prior_eval_kernel<<<grid, kPriorThreads, sm, h->stream>>>(*prior, residuals, jacobians, status);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

extern "C" isv_status isv_add_marg_prior(isv_handle* h, const isv_marg_prior* prior, const double* residuals,
                                         const isv_marg_generic_in* in, const isv_marg_generic_out* out, int32_t problem) {
  if (!h || !prior_ok(prior) || !residuals || !in || !out || !out->A || !out->b) return ISV_ERR_BAD_ARG;
  if (problem < -1 || problem >= in->n_problems || in->pos < 1 || in->n_problems > 65535) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  const int nt = (prior->n + kPaTile - 1) / kPaTile;
  const int first = problem < 0 ? 0 : problem, count = problem < 0 ? in->n_problems : 1;   // -1: every problem of the batch
  marg_prior_add_kernel<<<dim3(nt, nt, count), kPriorThreads, 0, h->stream>>>(
      *prior, residuals, out->A + (size_t)first * in->pos * in->pos, out->b + (size_t)first * in->pos, in->pos,
      out->status ? out->status + first : nullptr, in->m_dense, in->m_dense + in->m_diag);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

// tridiagonalization + QL (rotation log) + log applied to Z for `np` symmetric n x n matrices A (stride a_stride
// doubles, read-only); W: n x n scratch per problem (stride w_stride), Z: [np][n][n] receives the eigenvectors
// (columns, unsorted); eigenvalues + flags stay in the handle's eigensolver scratch (h->eig).
static isv_status sym_eig_launch(isv_handle* h, int n, int np, const double* A, size_t a_stride, double* W, size_t w_stride,
                                 double* Z) {
  if (n < 1 || n > kSeMaxN || np < 1 || np > 65535) return ISV_ERR_BAD_ARG;
  const size_t need = (size_t)np * sym_eig_scratch_bytes(n);
  if (h->eig_bytes < need) {
    if (h->eig) {
      ISV_CUDA(cudaStreamSynchronize(h->stream));
      ISV_CUDA(cudaFree(h->eig));
      h->eig = nullptr;
      h->eig_bytes = 0;
    }
    if (cudaMalloc(&h->eig, need) != cudaSuccess) {
      cudaGetLastError();
      return ISV_ERR_ALLOC;
    }
    h->eig_bytes = need;
  }
  const size_t sm1 = sym_tridiag_smem_doubles(n) * sizeof(double);
  ISV_CUDA(cudaFuncSetAttribute(sym_tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  // small matrices: fewer threads per CTA so that several problems share an SM
  const int t1 = n <= 64 ? 128 : (n <= 128 ? 256 : kSeThreads);
  sym_tridiag_kernel<<<np, t1, sm1, h->stream>>>(n, A, a_stride, W, w_stride, h->eig);
  const int rows = ql_apply_rows(n);
  const size_t sm3 = ql_apply_smem_bytes(n, rows);
  const dim3 slabs((n + rows - 1) / rows, np);
  ISV_CUDA(cudaFuncSetAttribute(q_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
  ISV_CUDA(cudaFuncSetAttribute(ql_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
  // fork: the rows of Q need only the reflectors, the serial QL only (d, e) -- they overlap on two streams
  ISV_CUDA(cudaEventRecord(h->aux_ev[0], h->stream));
  ISV_CUDA(cudaStreamWaitEvent(h->aux[0], h->aux_ev[0], 0));
  q_rows_kernel<<<slabs, kQaThreads, sm3, h->aux[0]>>>(n, rows, W, w_stride, Z, h->eig);
  ISV_CUDA(cudaEventRecord(h->aux_ev[1], h->aux[0]));
  tridiag_ql_kernel<<<np, 32, 2 * (size_t)n * sizeof(double), h->stream>>>(n, h->eig);
  ISV_CUDA(cudaStreamWaitEvent(h->stream, h->aux_ev[1], 0));
  ql_apply_kernel<<<slabs, kQaThreads, sm3, h->stream>>>(n, rows, Z, h->eig);
  h->launches += 4;
  ISV_CUDA(cudaGetLastError());
  return ISV_OK;
}

static isv_status schur_eig_impl(isv_handle* h, const isv_marg_generic_in* in, const isv_marg_generic_out* out,
                                 int schur_only) {
  ISV_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)(in->pos - in->m_dense - in->m_diag);
  // handle-owned scratch per problem: T = A_rm pinv (n x m_dense) of marg_schur_eig_kernel, later the eigenvector
  // buffer Z (n x n) of the reduced system -- sized for the larger of the two (m_dense may exceed n)
  const size_t gcols = n > (size_t)in->m_dense ? n : (size_t)in->m_dense;
  const size_t need = (size_t)in->n_problems * n * gcols * sizeof(double);
  if (h->gram_bytes < need) {
    if (h->gram) {
      ISV_CUDA(cudaStreamSynchronize(h->stream));
      ISV_CUDA(cudaFree(h->gram));
      h->gram = nullptr;
      h->gram_bytes = 0;
    }
    if (cudaMalloc(&h->gram, need + 256) != cudaSuccess) {
      cudaGetLastError();
      return ISV_ERR_ALLOC;
    }
    h->gram_bytes = need + 256;
  }
  const size_t sm = (4 * kMgMaxDense * kMgMaxDense + 6 * 16 + 32) * sizeof(double);
  ISV_CUDA(cudaFuncSetAttribute(marg_schur_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  if (in->m_diag > 0) {
    const int nt = (in->pos - in->m_diag + kSdTile - 1) / kSdTile;
    schur_diag_dmma_kernel<<<dim3(nt * (nt + 1) / 2, in->n_problems), kSdThreads, 0, h->stream>>>(*in, *out);
    ++h->launches;
  }
  marg_schur_eig_kernel<<<in->n_problems, kMgThreads, sm, h->stream>>>(*in, *out, h->gram, schur_only, 1);
  ++h->launches;
  ISV_CUDA(cudaGetLastError());
  if (!schur_only) {   // SelfAdjointEigenSolver of the reduced system -> linearized_jacobians / residuals
    // W scratch = the (consumed) normal equations of each problem, Z scratch = the handle's factor buffer
    isv_status st = sym_eig_launch(h, (int)n, in->n_problems, out->A_red, n * n, out->A, (size_t)in->pos * in->pos, h->gram);
    if (st != ISV_OK) return st;
    eig_prior_kernel<<<in->n_problems, 256, n * (sizeof(double) + sizeof(int)), h->stream>>>(
        (int)n, out->b_red, h->gram, h->eig, out->linearized_jacobians, out->linearized_residuals, out->rank, out->status,
        in->eps);
    ++h->launches;
    ISV_CUDA(cudaGetLastError());
  }
  return ISV_OK;
}

extern "C" isv_status isv_test_sym_eig(isv_handle* h, int nb, int n, const double* A, double* lam, double* V, int32_t* info) {
  if (!h || nb <= 0 || n < 1 || n > kSeMaxN || !A || !lam || !V || !info) return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  double *dA, *dW, *dZ, *dV, *dl;
  int* di;
  const size_t mb = sizeof(double) * (size_t)nb * n * n;
  ISV_CUDA(cudaMalloc(&dA, mb));
  ISV_CUDA(cudaMalloc(&dW, mb));
  ISV_CUDA(cudaMalloc(&dZ, mb));
  ISV_CUDA(cudaMalloc(&dV, mb));
  ISV_CUDA(cudaMalloc(&dl, sizeof(double) * (size_t)nb * n));
  ISV_CUDA(cudaMalloc(&di, sizeof(int) * 2 * (size_t)nb));
  ISV_CUDA(cudaMemcpyAsync(dA, A, mb, cudaMemcpyHostToDevice, h->stream));
  isv_status st = sym_eig_launch(h, n, nb, dA, (size_t)n * n, dW, (size_t)n * n, dZ);
  if (st == ISV_OK) {
    eig_test_out_kernel<<<nb, 256, n * sizeof(int), h->stream>>>(n, dZ, h->eig, dl, dV, di);
    ++h->launches;
    ISV_CUDA(cudaGetLastError());
    ISV_CUDA(cudaMemcpyAsync(V, dV, mb, cudaMemcpyDeviceToHost, h->stream));
    ISV_CUDA(cudaMemcpyAsync(lam, dl, sizeof(double) * (size_t)nb * n, cudaMemcpyDeviceToHost, h->stream));
    ISV_CUDA(cudaMemcpyAsync(info, di, sizeof(int) * 2 * (size_t)nb, cudaMemcpyDeviceToHost, h->stream));
    ISV_CUDA(cudaStreamSynchronize(h->stream));
  }
  cudaFree(dA); cudaFree(dW); cudaFree(dZ); cudaFree(dV); cudaFree(dl); cudaFree(di);
  return st;
}

// ---- MarginalizationInfo::preMarginalize + marginalize from host memory ---------------------------------------
extern "C" isv_status isv_marginalize_host(isv_handle* h, const isv_marg_host_in* in, isv_marg_host_out* out) {
  if (!h || !in || !out || !pb_ok(&in->pb)) return ISV_ERR_BAD_ARG;
  const isv_param_blocks& pb = in->pb;
  const isv_proj_factors& pf = in->proj;
  const isv_imu_factors& mf = in->imu;
  const isv_small_factors& sf = in->small_factors;
  if (pf.n < 0 || mf.n < 0 || sf.n_rel < 0 || sf.n_se3 < 0 || sf.n_vb < 0 || sf.n_rp < 0 || sf.n_yaw < 0) return ISV_ERR_BAD_ARG;
  const bool with_td = pf.td_obs != nullptr;
  if (with_td && (!pf.td || pf.n_td < 1 || !in->pos_td)) return ISV_ERR_BAD_ARG;
  const isv_marg_prior* prior = in->prior;
  if (prior && !prior_ok(prior)) return ISV_ERR_BAD_ARG;
  if ((pb.n_pose && !in->pos_pose) || (pb.n_speed_bias && !in->pos_speed_bias) || (pb.n_ex_pose && !in->pos_ex_pose) ||
      (pb.n_feature && !in->pos_feature))
    return ISV_ERR_BAD_ARG;
  const int n = in->pos - in->m_dense - in->m_diag;
  if (in->pos < 1 || n < 1 || in->m_dense < 0 || in->m_diag < 0 || in->m_dense > kMgMaxDense) return ISV_ERR_BAD_ARG;
  if (!out->A_red || !out->b_red || (!in->schur_only && (!out->linearized_jacobians || !out->linearized_residuals)))
    return ISV_ERR_BAD_ARG;
  ISV_CUDA(cudaSetDevice(h->device));
  const size_t D = sizeof(double), P = (size_t)pf.n, NI = (size_t)mf.n;
  bool ex_live = false;
  for (int e = 0; e < pb.n_ex_pose; ++e) ex_live = ex_live || in->pos_ex_pose[e] >= 0;
  // ---- `values`: every Evaluate output, contiguous (offsets in doubles) ----------------------------------
  size_t voff = 0;
  auto vcarve = [&](size_t cnt) { size_t o = voff; voff += cnt; return o; };
  const size_t v_pr = vcarve(2 * P), v_pi = vcarve(14 * P), v_pj = vcarve(14 * P), v_pe = vcarve(ex_live ? 14 * P : 0),
               v_pfe = vcarve(2 * P), v_pt = vcarve(with_td ? 2 * P : 0), v_ir = vcarve(15 * NI), v_ij = vcarve(ISV_IMU_JAC_REC * NI),
               v_rr = vcarve(6 * (size_t)sf.n_rel), v_rj = vcarve(84 * (size_t)sf.n_rel), v_sr = vcarve(6 * (size_t)sf.n_se3),
               v_sj = vcarve(42 * (size_t)sf.n_se3), v_vr = vcarve(9 * (size_t)sf.n_vb), v_vj = vcarve(81 * (size_t)sf.n_vb),
               v_qr = vcarve(2 * (size_t)sf.n_rp), v_qj = vcarve(14 * (size_t)sf.n_rp), v_yr = vcarve((size_t)sf.n_yaw),
               v_yj = vcarve(7 * (size_t)sf.n_yaw);
  // ---- block tables, built on the host from the factor index lists and the position tables ---------------
  std::vector<isv_ne_factor> facs;
  std::vector<isv_ne_block> blks;
  bool bad_index = false;
  auto add_block = [&](size_t jac, int stride, int ls, int position) {
    if (position < 0) return;
    blks.push_back(isv_ne_block{(int64_t)jac, stride, ls, position, 0});
  };
  auto chk = [&](int idx, int cnt) { if (idx < 0 || idx >= cnt) bad_index = true; return !(idx < 0 || idx >= cnt); };
  for (size_t k = 0; k < P; ++k) {
    const int i = pf.idx[k], j = pf.idx[pf.stride + k], e = pf.idx[2 * pf.stride + k], f = pf.idx[3 * pf.stride + k];
    if (!chk(i, pb.n_pose) || !chk(j, pb.n_pose) || !chk(e, pb.n_ex_pose) || !chk(f, pb.n_feature)) continue;
    const int first = (int)blks.size();
    add_block(v_pi + 14 * k, 7, 6, in->pos_pose[i]);
    add_block(v_pj + 14 * k, 7, 6, in->pos_pose[j]);
    if (ex_live) add_block(v_pe + 14 * k, 7, 6, in->pos_ex_pose[e]);
    add_block(v_pfe + 2 * k, 1, 1, in->pos_feature[f]);
    if (with_td) {
      const int t = pf.td_idx ? pf.td_idx[k] : 0;
      if (!chk(t, pf.n_td)) { blks.resize(first); continue; }
      add_block(v_pt + 2 * k, 1, 1, in->pos_td[t]);
    }
    if ((int)blks.size() > first) facs.push_back(isv_ne_factor{(int64_t)(v_pr + 2 * k), 2, (int)blks.size() - first, first, 0});
  }
  for (size_t k = 0; k < NI; ++k) {
    const int i = mf.idx[2 * k], j = mf.idx[2 * k + 1];
    if (!chk(i, pb.n_pose) || !chk(j, pb.n_pose) || !chk(i, pb.n_speed_bias) || !chk(j, pb.n_speed_bias)) continue;
    const int first = (int)blks.size();
    const size_t base = v_ij + ISV_IMU_JAC_REC * k;
    add_block(base, 7, 6, in->pos_pose[i]);
    add_block(base + 105, 9, 9, in->pos_speed_bias[i]);
    add_block(base + 240, 7, 6, in->pos_pose[j]);
    add_block(base + 345, 9, 9, in->pos_speed_bias[j]);
    if ((int)blks.size() > first) facs.push_back(isv_ne_factor{(int64_t)(v_ir + 15 * k), 15, (int)blks.size() - first, first, 0});
  }
  for (int k = 0; k < sf.n_rel; ++k) {
    const int i = sf.rel_idx[2 * k], j = sf.rel_idx[2 * k + 1];
    if (!chk(i, pb.n_pose) || !chk(j, pb.n_pose)) continue;
    const int first = (int)blks.size();
    add_block(v_rj + 84 * (size_t)k, 7, 6, in->pos_pose[i]);
    add_block(v_rj + 84 * (size_t)k + 42, 7, 6, in->pos_pose[j]);
    if ((int)blks.size() > first) facs.push_back(isv_ne_factor{(int64_t)(v_rr + 6 * (size_t)k), 6, (int)blks.size() - first, first, 0});
  }
  auto single = [&](int cnt, const int32_t* idx, const int32_t* postab, int n_blocks_family, size_t vres, size_t vjac, int nres,
                    int width, int stride, int ls) {
    for (int k = 0; k < cnt; ++k) {
      if (!chk(idx[k], n_blocks_family)) continue;
      const int first = (int)blks.size();
      add_block(vjac + (size_t)width * k, stride, ls, postab[idx[k]]);
      if ((int)blks.size() > first) facs.push_back(isv_ne_factor{(int64_t)(vres + (size_t)nres * k), nres, 1, first, 0});
    }
  };
  single(sf.n_se3, sf.se3_idx, in->pos_pose, pb.n_pose, v_sr, v_sj, 6, 42, 7, 6);
  single(sf.n_vb, sf.vb_idx, in->pos_speed_bias, pb.n_speed_bias, v_vr, v_vj, 9, 81, 9, 9);
  single(sf.n_rp, sf.rp_idx, in->pos_pose, pb.n_pose, v_qr, v_qj, 2, 14, 7, 6);
  single(sf.n_yaw, sf.yaw_idx, in->pos_pose, pb.n_pose, v_yr, v_yj, 1, 7, 7, 6);
  size_t prior_x = 0;   // doubles in the prior's x / x0
  if (prior) {
    for (int k = 0; k < prior->n_blocks; ++k) {
      const isv_prior_block& bl = prior->blocks[k];
      const int ls = bl.global_size == 7 ? 6 : bl.global_size;
      if (bl.global_size < 1 || bl.idx < 0 || bl.idx + ls > prior->n || bl.x_offset < 0 || bl.pos + ls > in->pos) bad_index = true;
      prior_x = std::max(prior_x, (size_t)bl.x_offset + (size_t)std::max(bl.global_size, 0));
    }
  }
  if (bad_index) { out->status = ISV_W_BAD_INDEX; return ISV_ERR_BAD_ARG; }
  // ---- device mirror ---------------------------------------------------------------------------------------
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  const size_t pos = (size_t)in->pos;
  // layout: [ every host input, one contiguous upload ] [ device-only ] [ every result, one contiguous download ]
  const size_t pn = prior ? (size_t)prior->n : 0, pnb = prior ? (size_t)prior->n_blocks : 0;
  const size_t o_pose = carve(pb.n_pose * 7 * D), o_sb = carve(pb.n_speed_bias * 9 * D), o_ex = carve(pb.n_ex_pose * 7 * D),
               o_ft = carve(pb.n_feature * D), o_pidx = carve(4 * P * 4), o_pobs = carve(5 * P * D), o_iidx = carve(2 * NI * 4),
               o_ipre = carve(NI * ISV_PREINT_REC * D), o_ridx = carve(2 * (size_t)sf.n_rel * 4),
               o_rrec = carve((size_t)sf.n_rel * ISV_REL_REC * D), o_sidx = carve((size_t)sf.n_se3 * 4),
               o_srec = carve((size_t)sf.n_se3 * ISV_SE3_REC * D), o_vidx = carve((size_t)sf.n_vb * 4),
               o_vrec = carve((size_t)sf.n_vb * ISV_VB_REC * D), o_qidx = carve((size_t)sf.n_rp * 4),
               o_qrec = carve((size_t)sf.n_rp * ISV_RP_REC * D), o_yidx = carve((size_t)sf.n_yaw * 4),
               o_yrec = carve((size_t)sf.n_yaw * ISV_YAW_REC * D),
               o_fac = carve(facs.size() * sizeof(isv_ne_factor)), o_blk = carve(blks.size() * sizeof(isv_ne_block)),
               o_tdo = carve(with_td ? 8 * P * D : 0), o_td = carve(with_td ? (size_t)pf.n_td * D : 0),
               o_tdi = carve(with_td && pf.td_idx ? P * 4 : 0),
               o_pblk = carve(pnb * sizeof(isv_prior_block)), o_pJ = carve(pn * pn * D), o_pr0 = carve(pn * D),
               o_px0 = carve(prior_x * D), o_px = carve(prior_x * D);
  const size_t in_bytes = off;
  const size_t o_val = carve(voff * D), o_A = carve(pos * pos * D), o_b = carve(pos * D), o_pres = carve(pn * D);
  const size_t out_begin = off;
  const size_t o_Ar = carve((size_t)n * n * D), o_br = carve(n * D), o_J = carve((size_t)n * n * D), o_r = carve(n * D),
               o_rank = carve(8), o_st = carve(8);
  const size_t out_bytes = off - out_begin;
  isv_status st = ensure_dbuf(h, off);
  if (st != ISV_OK) return st;
  st = ensure_pinned(h, in_bytes + out_bytes);
  if (st != ISV_OK) return st;
  char* d = h->dbuf;
  char* hp = h->pinned;
  cudaStream_t s = h->stream;
  // the ~25 small arrays are packed into the pinned block and go up in ONE copy (each pageable cudaMemcpyAsync costs
  // ~10 us of staging: a quarter of the call at VINS problem sizes)
  auto up = [&](size_t o, const void* src, size_t bytes) {
    if (bytes) memcpy(hp + o, src, bytes);
    return cudaSuccess;
  };
  ISV_CUDA(up(o_pose, pb.pose, pb.n_pose * 7 * D));
  ISV_CUDA(up(o_sb, pb.speed_bias, pb.n_speed_bias * 9 * D));
  ISV_CUDA(up(o_ex, pb.ex_pose, pb.n_ex_pose * 7 * D));
  ISV_CUDA(up(o_ft, pb.feature, pb.n_feature * D));
  for (int c = 0; c < 4 && P; ++c) ISV_CUDA(up(o_pidx + c * P * 4, pf.idx + c * pf.stride, P * 4));
  for (int c = 0; c < 5 && P; ++c) ISV_CUDA(up(o_pobs + c * P * D, pf.obs + c * pf.stride, P * D));
  ISV_CUDA(up(o_iidx, mf.idx, 2 * NI * 4));
  ISV_CUDA(up(o_ipre, mf.preint, NI * ISV_PREINT_REC * D));
  ISV_CUDA(up(o_ridx, sf.rel_idx, 2 * (size_t)sf.n_rel * 4));
  ISV_CUDA(up(o_rrec, sf.rel_rec, (size_t)sf.n_rel * ISV_REL_REC * D));
  ISV_CUDA(up(o_sidx, sf.se3_idx, (size_t)sf.n_se3 * 4));
  ISV_CUDA(up(o_srec, sf.se3_rec, (size_t)sf.n_se3 * ISV_SE3_REC * D));
  ISV_CUDA(up(o_vidx, sf.vb_idx, (size_t)sf.n_vb * 4));
  ISV_CUDA(up(o_vrec, sf.vb_rec, (size_t)sf.n_vb * ISV_VB_REC * D));
  ISV_CUDA(up(o_qidx, sf.rp_idx, (size_t)sf.n_rp * 4));
  ISV_CUDA(up(o_qrec, sf.rp_rec, (size_t)sf.n_rp * ISV_RP_REC * D));
  ISV_CUDA(up(o_yidx, sf.yaw_idx, (size_t)sf.n_yaw * 4));
  ISV_CUDA(up(o_yrec, sf.yaw_rec, (size_t)sf.n_yaw * ISV_YAW_REC * D));
  if (with_td) {
    for (int c = 0; c < 8 && P; ++c) ISV_CUDA(up(o_tdo + c * P * D, pf.td_obs + c * pf.stride, P * D));
    ISV_CUDA(up(o_td, pf.td, (size_t)pf.n_td * D));
    if (pf.td_idx) ISV_CUDA(up(o_tdi, pf.td_idx, P * 4));
  }
  if (prior) {
    ISV_CUDA(up(o_pblk, prior->blocks, pnb * sizeof(isv_prior_block)));
    ISV_CUDA(up(o_pJ, prior->linearized_jacobians, pn * pn * D));
    ISV_CUDA(up(o_pr0, prior->linearized_residuals, pn * D));
    ISV_CUDA(up(o_px0, prior->x0, prior_x * D));
    ISV_CUDA(up(o_px, prior->x, prior_x * D));
  }
  ISV_CUDA(up(o_fac, facs.data(), facs.size() * sizeof(isv_ne_factor)));
  ISV_CUDA(up(o_blk, blks.data(), blks.size() * sizeof(isv_ne_block)));
  ISV_CUDA(cudaMemcpyAsync(d, hp, in_bytes, cudaMemcpyHostToDevice, s));
  ISV_CUDA(cudaMemsetAsync(d + o_st, 0, 8, s));
  // ---- preMarginalize: Evaluate every residual block -------------------------------------------------------
  double* V = (double*)(d + o_val);
  isv_param_blocks dpb = {pb.n_pose, pb.n_speed_bias, pb.n_ex_pose, pb.n_feature, (const double*)(d + o_pose),
                          (const double*)(d + o_sb), (const double*)(d + o_ex), (const double*)(d + o_ft)};
  int32_t* dst = (int32_t*)(d + o_st);
  if (P) {
    isv_proj_factors dpf = pf;
    dpf.stride = (int64_t)P;
    dpf.idx = (const int32_t*)(d + o_pidx);
    dpf.obs = (const double*)(d + o_pobs);
    if (with_td) {
      dpf.td_obs = (const double*)(d + o_tdo);
      dpf.td = (const double*)(d + o_td);
      dpf.td_idx = pf.td_idx ? (const int32_t*)(d + o_tdi) : nullptr;
    }
    isv_proj_eval po = {V + v_pr, V + v_pi, V + v_pj, ex_live ? V + v_pe : nullptr, V + v_pfe, with_td ? V + v_pt : nullptr};
    st = eval_projection_on(h, s, &dpb, &dpf, &po, dst);
    if (st != ISV_OK) return st;
  }
  if (NI) {
    isv_imu_factors dmf = {mf.n, (const int32_t*)(d + o_iidx), (const double*)(d + o_ipre)};
    isv_imu_eval mo = {V + v_ir, V + v_ij};
    st = eval_imu_on(h, s, &dpb, &dmf, &mo, dst);
    if (st != ISV_OK) return st;
  }
  if (sf.n_rel + sf.n_se3 + sf.n_vb + sf.n_rp + sf.n_yaw > 0) {
    isv_small_factors dsf = sf;
    dsf.rel_idx = (const int32_t*)(d + o_ridx); dsf.rel_rec = (const double*)(d + o_rrec);
    dsf.se3_idx = (const int32_t*)(d + o_sidx); dsf.se3_rec = (const double*)(d + o_srec);
    dsf.vb_idx = (const int32_t*)(d + o_vidx); dsf.vb_rec = (const double*)(d + o_vrec);
    dsf.rp_idx = (const int32_t*)(d + o_qidx); dsf.rp_rec = (const double*)(d + o_qrec);
    dsf.yaw_idx = (const int32_t*)(d + o_yidx); dsf.yaw_rec = (const double*)(d + o_yrec);
    isv_small_eval so = {V + v_rr, V + v_rj, V + v_sr, V + v_sj, V + v_vr, V + v_vj, V + v_qr, V + v_qj, V + v_yr, V + v_yj};
    st = eval_small_on(h, s, &dpb, &dsf, &so, dst);
    if (st != ISV_OK) return st;
  }
  // ---- marginalize ---------------------------------------------------------------------------------------------
  isv_marg_generic_in gi = {1, in->pos, in->m_dense, in->m_diag, (int64_t)facs.size(), (const isv_ne_factor*)(d + o_fac),
                            (const isv_ne_block*)(d + o_blk), V, in->eps};
  isv_marg_generic_out go = {(double*)(d + o_A), (double*)(d + o_b), (double*)(d + o_Ar), (double*)(d + o_br),
                             (double*)(d + o_J), (double*)(d + o_r), (int32_t*)(d + o_rank), dst + 1};
  ISV_CUDA(cudaMemsetAsync(d + o_rank, 0, 8, s));
  // (isv_build_normal_equations zeroes status[problem]: give it its own word, merged below)
  if (!prior) {
    st = marginalize_generic_impl(h, &gi, &go, in->schur_only ? 1 : 0);
  } else {   // MarginalizationFactor::Evaluate, then its J^T J / J^T r on top of the ordinary blocks' normal equations
    isv_marg_prior dp = *prior;
    dp.blocks = (const isv_prior_block*)(d + o_pblk);
    dp.linearized_jacobians = (const double*)(d + o_pJ);
    dp.linearized_residuals = (const double*)(d + o_pr0);
    dp.x0 = (const double*)(d + o_px0);
    dp.x = (const double*)(d + o_px);
    st = marg_generic_check(h, &gi, &go, true);
    if (st == ISV_OK) st = isv_eval_marg_prior(h, &dp, (double*)(d + o_pres), nullptr, dst);
    if (st == ISV_OK) st = isv_build_normal_equations(h, &gi, &go);
    if (st == ISV_OK) st = isv_add_marg_prior(h, &dp, (const double*)(d + o_pres), &gi, &go, 0);
    if (st == ISV_OK) st = schur_eig_impl(h, &gi, &go, in->schur_only ? 1 : 0);
  }
  if (st != ISV_OK) return st;
  char* ho = hp + in_bytes;   // results come down in one copy into the pinned block
  ISV_CUDA(cudaMemcpyAsync(ho, d + out_begin, out_bytes, cudaMemcpyDeviceToHost, s));
  ISV_CUDA(cudaStreamSynchronize(s));
  memcpy(out->A_red, ho + (o_Ar - out_begin), (size_t)n * n * D);
  memcpy(out->b_red, ho + (o_br - out_begin), n * D);
  if (!in->schur_only) {
    memcpy(out->linearized_jacobians, ho + (o_J - out_begin), (size_t)n * n * D);
    memcpy(out->linearized_residuals, ho + (o_r - out_begin), n * D);
  }
  int32_t hst[2], hrank;
  memcpy(hst, ho + (o_st - out_begin), 8);
  memcpy(&hrank, ho + (o_rank - out_begin), 4);
  out->status = hst[0] | hst[1];
  out->rank = hrank;
  return ISV_OK;
}
