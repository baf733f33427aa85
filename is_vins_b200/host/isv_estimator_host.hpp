// C++ host side of the drop-in boundary: the member functions of the reference's `Estimator` that make
// up the hot path, with the reference's names, members and ownership rules, implemented as
// pack -> C ABI (include/isv_capi.h) -> unpack.  No arithmetic of the path happens here and there is no
// CPU fallback: without libisv_b200.so / an sm_100 GPU every method throws.
//
//   Estimator::initFactorGraph()  sparsification tail   /root/reference/src/estimator.cpp:745-1001
//   Estimator::MargForward()                            /root/reference/src/estimator.cpp:1149-1352
//   Estimator::MargBackward()                           /root/reference/src/estimator.cpp:1354-1539
//   Estimator::slideWindow()      factor rotation       /root/reference/src/estimator.cpp:1605-1638
//   IntegrationBase::push_back / repropagate            /root/reference/include/factor/integration_base.h:30-52
//
// Eigen is not available in this build, so the factor classes below carry plain arrays with exactly
// the bytes of the reference's Eigen members (column-major): `sqrt_info` of a SE3PriorFactor is the
// 36 doubles of `Eigen::MatrixXd sqrt_info` (6x6), `R` the 9 doubles of `Eigen::Matrix3d R`, ... so a
// build against the real classes replaces `std::memcpy(dst, src, n)` by `m.data()`.
#pragma once
#include <algorithm>
#include <cstring>
#include <queue>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/isv_capi.h"

namespace isv_host {

constexpr int SIZE_POSE = 7, SIZE_SPEEDBIAS = 9, SIZE_FEATURE = 1;   // include/parameters.h:83-87

// ---- factor classes: members named as in include/factor/*.h ------------------------------------
struct SE3PriorFactor {            // se3_prior_factor.h:9-141
  double t[3], R[9], sqrt_info[36];
  int index = -1;
  void setIndex(int i) { index = i; }
};
struct RelativePoseFactor {        // relative_pose_factor.h:13-196
  double delta_t[3], delta_R[9], sqrt_info[36];
  int imu_i = -1, imu_j = -1;
  void setIndex(int i, int j) { imu_i = i; imu_j = j; }
  void shift() { --imu_i; --imu_j; }                                  // :126-129
};
struct Linear9Factor {             // linear9_factor.h:8-74
  double VB[9], sqrt_info[81];
  int index = -1;
  void setIndex(int i) { index = i; }
};
struct RollPitchFactor {           // rollpitch_factor.h:10-137
  double R[9], sqrt_info[4];
  int index = -1;
  void setIndex(int i) { index = i; }
  void shift() { --index; }
};
struct ProjectionFactor {          // projection_factor.h:9-55 (members read by MargForward)
  double pts_i[3], pts_j[3];
  int imu_i = -1, imu_j = -1, feature_idx = -1;
  void setIndex(int i, int j, int f) { imu_i = i; imu_j = j; feature_idx = f; }
};
struct CombinedFactors {           // pose_graph_factors.h:6-52
  RelativePoseFactor* relativePoseFactor = nullptr;
  RollPitchFactor* rollPitchFactor = nullptr;
  long vio_index = -1;
  int length = 0;
  long pg_index = 0;
  double covRel[36] = {0}, covAbs[4] = {0};
  double distance = 0, ts = 0;
  double Ri[9] = {0}, ti[3] = {0};
};

// IntegrationBase (integration_base.h:13-208): the host keeps the raw sample buffers, the
// propagation itself runs on the GPU (isv_preintegrate_host) when the record is needed.
struct IntegrationBase {
  double acc_0[3], gyr_0[3], linearized_ba[3], linearized_bg[3];
  std::vector<double> dt_buf, acc_buf, gyr_buf;    // acc_buf / gyr_buf: 3 doubles per sample
  double record[ISV_PREINT_REC];                   // delta_p, delta_q, delta_v, biases, sum_dt, jacobian, covariance
  bool dirty = true;
  IntegrationBase(const double* a0, const double* g0, const double* ba, const double* bg) {
    std::memcpy(acc_0, a0, 24); std::memcpy(gyr_0, g0, 24);
    std::memcpy(linearized_ba, ba, 24); std::memcpy(linearized_bg, bg, 24);
  }
  void push_back(double dt, const double* acc, const double* gyr) {   // :30-36
    dt_buf.push_back(dt);
    acc_buf.insert(acc_buf.end(), acc, acc + 3);
    gyr_buf.insert(gyr_buf.end(), gyr, gyr + 3);
    dirty = true;
  }
  void repropagate(const double* ba, const double* bg) {              // :38-52
    std::memcpy(linearized_ba, ba, 24); std::memcpy(linearized_bg, bg, 24);
    dirty = true;
  }
  double sum_dt() const { return record[16]; }
};
struct IMUFactor {                 // imu_factor.h:13-274
  IntegrationBase* pre_integration;
  int imu_i = -1, imu_j = -1;
  explicit IMUFactor(IntegrationBase* p) : pre_integration(p) {}
  void setIndex(int i, int j) { imu_i = i; imu_j = j; }
};

inline void check(isv_status st, const char* what) {
  if (st != ISV_OK) throw std::runtime_error(std::string(what) + ": " + isv_status_string(st));
}

// ---- the hot-path slice of `class Estimator` (include/estimator.h:22-160) ------------------------
class Estimator {
 public:
  static constexpr int Vo_SIZE = 8, ALL_BUF_SIZE = 18, NUM_OF_F = 1000;    // include/parameters.h:35-40
  double para_Pose[ALL_BUF_SIZE][SIZE_POSE];
  double para_SpeedBias[ALL_BUF_SIZE][SIZE_SPEEDBIAS];
  double para_Feature[NUM_OF_F][SIZE_FEATURE];
  double para_Ex_Pose[1][SIZE_POSE];
  double Headers[ALL_BUF_SIZE] = {0};
  double Ps[ALL_BUF_SIZE][3] = {{0}}, Rs[ALL_BUF_SIZE][9] = {{0}};
  IntegrationBase* pre_integrations[ALL_BUF_SIZE] = {nullptr};
  long PoseGraphFactorCount = 0;
  std::queue<CombinedFactors*> pose_graph_factors_buf;
  // system factors
  Linear9Factor* vioVBPrior = nullptr;
  std::vector<RelativePoseFactor*> vioRelativePoseEdges;
  std::vector<RollPitchFactor*> vioRollPitchEdges;
  SE3PriorFactor* vioPosePriorEdge = nullptr;
  // forward prior / result
  std::vector<int> MargPointIdx;
  std::vector<ProjectionFactor*> forwardProjectiontoSparsify;
  SE3PriorFactor* forwardPosePriorEdgeToAdd = nullptr;
  // backward prior / result
  IMUFactor* backwardIMUtoSparsify = nullptr;
  Linear9Factor* backwardVBEdgeToAdd = nullptr;
  RelativePoseFactor* backwardRelativePoseEdgeToAdd = nullptr;
  // diagnostics the reference does not have (void + assert): ranks and ISV_W_* status bits
  int last_fwd_rank = 0, last_bwd_rank = 0, last_init_rank = 0, last_status = 0;

  explicit Estimator(int device = 0, const isv_config* cfg = nullptr) {
    isv_config c;
    isv_default_config(&c);
    if (cfg) c = *cfg;
    c.vo_size = Vo_SIZE;
    c.all_buf_size = ALL_BUF_SIZE;
    check(isv_create(&c, device, &h_), "isv_create");
  }
  ~Estimator() { isv_destroy(h_); }
  Estimator(const Estimator&) = delete;
  Estimator& operator=(const Estimator&) = delete;

  // pre-integration record of one interval, computed on the GPU on demand
  const double* preintegrated(IntegrationBase* p) {
    if (p->dirty) {
      const int K = (int)p->dt_buf.size();
      std::vector<double> raw((size_t)K * 7);
      for (int k = 0; k < K; ++k) {
        raw[7 * k] = p->dt_buf[k];
        std::memcpy(&raw[7 * k + 1], &p->acc_buf[3 * k], 24);
        std::memcpy(&raw[7 * k + 4], &p->gyr_buf[3 * k], 24);
      }
      double init[12];
      std::memcpy(init, p->acc_0, 24); std::memcpy(init + 3, p->gyr_0, 24);
      std::memcpy(init + 6, p->linearized_ba, 24); std::memcpy(init + 9, p->linearized_bg, 24);
      isv_preint_in in{1, K, nullptr, raw.data(), init};
      check(isv_preintegrate_host(h_, &in, p->record), "isv_preintegrate_host");
      p->dirty = false;
    }
    return p->record;
  }

  // initFactorGraph(), sparsification tail (:745-1001): V-1 IMU factors -> vioRelativePoseEdges[1..V-1],
  // vioPosePriorEdge, vioVBPrior.  (The ceres solve before it is out of scope.)
  void initFactorGraph() {
    std::vector<double> pre((size_t)(Vo_SIZE - 1) * ISV_PREINT_REC);
    for (int i = 0; i < Vo_SIZE - 1; ++i)
      std::memcpy(&pre[(size_t)i * ISV_PREINT_REC], preintegrated(pre_integrations[i + 1]), sizeof(double) * ISV_PREINT_REC);
    std::vector<double> rel((size_t)(Vo_SIZE - 1) * ISV_REL_REC);
    double se3[ISV_SE3_REC], vb[ISV_VB_REC];
    int32_t rank = 0, status = 0;
    isv_init_in in{1, &para_Pose[0][0], &para_SpeedBias[0][0], pre.data()};
    isv_init_out out{rel.data(), se3, vb, &rank, &status};
    check(isv_init_sparsify_host(h_, &in, &out), "isv_init_sparsify_host");
    last_init_rank = rank;
    last_status = status;
    vioRelativePoseEdges.assign(Vo_SIZE, nullptr);                     // [0] is nullptr (:821)
    for (int i = 0; i < Vo_SIZE - 1; ++i) {
      auto* f = new RelativePoseFactor();
      unpack48(&rel[(size_t)i * ISV_REL_REC], f->delta_t, f->delta_R, f->sqrt_info);
      f->setIndex(i, i + 1);
      vioRelativePoseEdges[i + 1] = f;
    }
    auto* p = new SE3PriorFactor();
    unpack48(se3, p->t, p->R, p->sqrt_info);
    p->setIndex(0);
    vioPosePriorEdge = p;
    auto* v = new Linear9Factor();
    std::memcpy(v->VB, vb, 72); std::memcpy(v->sqrt_info, vb + 9, 648);
    v->setIndex(Vo_SIZE - 1);
    vioVBPrior = v;
  }

  void MargForward() {                                                // :1149-1352
    FwdPack p;
    packForward(p);
    isv_fwd_out out;
    check(isv_marg_forward(h_, &p.in, &out), "isv_marg_forward");
    last_status = 0;
    unpackForward(out, p.rp_valid);
  }

  void MargBackward() {                                               // :1354-1539
    BwdPack p;
    packBackward(p);
    isv_bwd_out out;
    check(isv_marg_backward(h_, &p.in, &out), "isv_marg_backward");
    unpackBackward(out);
  }

  // `MargForward(); MargBackward();` of one MARGIN_OLD event (:1555-1558) in ONE blocking C-ABI call: the two read
  // disjoint members and neither reads the other's outputs, so their kernel chains run forked on two streams behind a
  // single host-to-device / device-to-host round trip.  Same members written, bit-identical values.
  void MargForwardBackward() {
    FwdPack pf;
    BwdPack pb;
    packForward(pf);
    packBackward(pb);
    isv_fwd_out fo;
    isv_bwd_out bo;
    check(isv_marg_event(h_, &pf.in, &pb.in, &fo, &bo), "isv_marg_event");
    last_status = 0;
    unpackForward(fo, pf.rp_valid);
    unpackBackward(bo);
  }

  // the factor rotation inside slideWindow() (:1605-1638); ownership exactly as the reference
  void slideWindowFactors() {
    for (int i = 1; i < Vo_SIZE; ++i) vioRelativePoseEdges[i]->shift();
    for (int i = 1; i < Vo_SIZE - 1; ++i) std::swap(vioRelativePoseEdges[i], vioRelativePoseEdges[i + 1]);
    for (auto* f : forwardProjectiontoSparsify) delete f;
    forwardProjectiontoSparsify.clear();
    MargPointIdx.clear();
    for (auto it = vioRollPitchEdges.begin(); it != vioRollPitchEdges.end();) {
      (*it)->shift();
      if ((*it)->index < 0) it = vioRollPitchEdges.erase(it);          // not deleted: used in the pose graph
      else ++it;
    }
    backwardRelativePoseEdgeToAdd->setIndex(Vo_SIZE - 2, Vo_SIZE - 1);
    vioRelativePoseEdges[Vo_SIZE - 1] = backwardRelativePoseEdgeToAdd;  // old one not deleted (:1627)
    delete vioPosePriorEdge;
    forwardPosePriorEdgeToAdd->setIndex(0);
    vioPosePriorEdge = forwardPosePriorEdgeToAdd;
    delete vioVBPrior;
    backwardVBEdgeToAdd->setIndex(Vo_SIZE - 1);
    vioVBPrior = backwardVBEdgeToAdd;
  }

  isv_handle* handle() { return h_; }

 private:
  struct FwdPack {
    std::vector<double> inv_dep, pts_i, pts_j;
    double pse3[ISV_SE3_REC], prel[ISV_REL_REC], prp[ISV_RP_IN_REC];
    bool rp_valid;
    isv_fwd_in in;
  };
  struct BwdPack {
    double pvb[ISV_VB_REC];
    isv_bwd_in in;
  };
  void packForward(FwdPack& p) {
    const int L = (int)MargPointIdx.size();
    p.inv_dep.resize(L); p.pts_i.resize(3 * (size_t)L); p.pts_j.resize(3 * (size_t)L);
    for (int k = 0; k < L; ++k) {      // order == MargPointIdx == OrderMap (:1159-1162)
      p.inv_dep[k] = para_Feature[MargPointIdx[k]][0];
      std::memcpy(&p.pts_i[3 * (size_t)k], forwardProjectiontoSparsify[k]->pts_i, 24);
      std::memcpy(&p.pts_j[3 * (size_t)k], forwardProjectiontoSparsify[k]->pts_j, 24);
    }
    std::memset(p.prp, 0, sizeof(p.prp));
    pack48(vioPosePriorEdge->t, vioPosePriorEdge->R, vioPosePriorEdge->sqrt_info, p.pse3);
    const RelativePoseFactor* r1 = vioRelativePoseEdges[1];
    pack48(r1->delta_t, r1->delta_R, r1->sqrt_info, p.prel);
    p.rp_valid = !vioRollPitchEdges.empty() && vioRollPitchEdges[0]->index == 0;   // :1265-1271
    if (p.rp_valid) { p.prp[0] = 1.0; std::memcpy(p.prp + 1, vioRollPitchEdges[0]->sqrt_info, 32); }
    p.in = isv_fwd_in{L, para_Pose[0], para_Pose[1], para_Ex_Pose[0], p.inv_dep.data(), p.pts_i.data(), p.pts_j.data(),
                      p.pse3, p.prel, p.prp};
  }
  void unpackForward(const isv_fwd_out& out, bool rp_valid) {
    last_fwd_rank = out.rank;
    last_status |= out.status;
    auto* se3 = new SE3PriorFactor();                                  // forwardPosePriorEdgeToAdd (:1291-1351)
    unpack48(out.se3, se3->t, se3->R, se3->sqrt_info);
    forwardPosePriorEdgeToAdd = se3;
    auto* pg = new RelativePoseFactor();                               // CombinedFactors (:1243-1283)
    unpack48(out.pg, pg->delta_t, pg->delta_R, pg->sqrt_info);
    auto* cmb = new CombinedFactors();
    cmb->relativePoseFactor = pg;
    cmb->rollPitchFactor = rp_valid ? vioRollPitchEdges[0] : nullptr;
    if (rp_valid) std::memcpy(cmb->covAbs, out.pg + 85, 32);
    std::memcpy(cmb->covRel, out.pg + 48, 288);
    cmb->distance = out.pg[84];
    cmb->vio_index = PoseGraphFactorCount++;
    cmb->ts = Headers[0];
    std::memcpy(cmb->Ri, Rs[0], 72); std::memcpy(cmb->ti, Ps[0], 24);
    pose_graph_factors_buf.push(cmb);
  }
  void packBackward(BwdPack& p) {
    std::memcpy(p.pvb, vioVBPrior->VB, 72); std::memcpy(p.pvb + 9, vioVBPrior->sqrt_info, 648);
    const double* pre = preintegrated(backwardIMUtoSparsify->pre_integration);
    p.in = isv_bwd_in{para_Pose[Vo_SIZE - 1], para_SpeedBias[Vo_SIZE - 1], para_Pose[Vo_SIZE], para_SpeedBias[Vo_SIZE], p.pvb, pre};
  }
  void unpackBackward(const isv_bwd_out& out) {
    last_bwd_rank = out.rank;
    last_status |= out.status;
    auto* rel = new RelativePoseFactor();
    unpack48(out.rel, rel->delta_t, rel->delta_R, rel->sqrt_info);
    auto* vbp = new Linear9Factor();
    std::memcpy(vbp->VB, out.vb, 72); std::memcpy(vbp->sqrt_info, out.vb + 9, 648);
    auto* rp = new RollPitchFactor();
    std::memcpy(rp->R, out.rp, 72); std::memcpy(rp->sqrt_info, out.rp + 9, 32);
    rp->setIndex(Vo_SIZE - 1);                                         // :1516
    vioRollPitchEdges.push_back(rp);                                   // :1536-1538
    backwardVBEdgeToAdd = vbp;
    backwardRelativePoseEdgeToAdd = rel;
  }
  static void pack48(const double* t, const double* R, const double* s, double* rec) {
    std::memcpy(rec, t, 24); std::memcpy(rec + 3, R, 72); std::memcpy(rec + 12, s, 288);
  }
  static void unpack48(const double* rec, double* t, double* R, double* s) {
    std::memcpy(t, rec, 24); std::memcpy(R, rec + 3, 72); std::memcpy(s, rec + 12, 288);
  }
  isv_handle* h_ = nullptr;
};

}  // namespace isv_host
