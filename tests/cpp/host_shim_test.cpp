// Drives isv_host::Estimator (is_vins_b200/host/isv_estimator_host.hpp) -- the C++ host side of the
// drop-in boundary -- through initFactorGraph() and R MARGIN_OLD frames exactly the way
// Estimator::backendOptimization() / slideWindow() do (/root/reference/src/estimator.cpp:1541-1562,
// :1605-1638), on a fixture written by tests/test_host_cpp_gpu.py from the oracle's chain, and checks
// every recovered factor against the oracle's records (relative 1e-9, ranks exact).
//   usage: host_shim_test <fixture.bin>      (little-endian doubles, see the Python writer)
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../is_vins_b200/host/isv_estimator_host.hpp"

using namespace isv_host;

static std::vector<double> g_data;
static size_t g_pos = 0;
static double rd() {
  if (g_pos >= g_data.size()) { fprintf(stderr, "fixture underrun\n"); exit(2); }
  return g_data[g_pos++];
}
static void rdv(double* dst, int n) { for (int i = 0; i < n; ++i) dst[i] = rd(); }
static double g_worst = 0.0;
static int g_fail = 0;

static void cmp(const char* what, int round, const double* got, int n) {
  double num = 0, den = 0;
  for (int i = 0; i < n; ++i) {
    const double ref = rd();
    num += (got[i] - ref) * (got[i] - ref);
    den += ref * ref;
  }
  const double e = den > 0 ? std::sqrt(num / den) : std::sqrt(num);
  if (e > g_worst) g_worst = e;
  if (!(e <= 1e-9)) { ++g_fail; fprintf(stderr, "MISMATCH round %d %s: rel err %.3e\n", round, what, e); }
}
static void cmp48(const char* what, int round, const double* t, const double* R, const double* s) {
  char b[64];
  snprintf(b, sizeof b, "%s.t", what); cmp(b, round, t, 3);
  snprintf(b, sizeof b, "%s.R", what); cmp(b, round, R, 9);
  snprintf(b, sizeof b, "%s.sqrt_info", what); cmp(b, round, s, 36);
}
static void expect_int(const char* what, int round, int got) {
  const int ref = (int)rd();
  if (got != ref) { ++g_fail; fprintf(stderr, "MISMATCH round %d %s: %d != %d\n", round, what, got, ref); }
}
static IntegrationBase* read_preintegration() {
  const int K = (int)rd();
  double a0[3], g0[3], ba[3], bg[3];
  rdv(a0, 3); rdv(g0, 3); rdv(ba, 3); rdv(bg, 3);
  auto* p = new IntegrationBase(a0, g0, ba, bg);
  for (int k = 0; k < K; ++k) {
    double s[7];
    rdv(s, 7);
    p->push_back(s[0], s + 1, s + 4);
  }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s fixture.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("fixture"); return 2; }
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  g_data.resize(bytes / 8);
  if (fread(g_data.data(), 8, g_data.size(), f) != g_data.size()) { fprintf(stderr, "short read\n"); return 2; }
  fclose(f);

  const int R = (int)rd(), V = (int)rd();
  if (V != Estimator::Vo_SIZE) { fprintf(stderr, "fixture V=%d\n", V); return 2; }
  Estimator est(0);

  // ---- initFactorGraph ------------------------------------------------------------------------------
  for (int i = 0; i < V; ++i) rdv(est.para_Pose[i], 7);
  for (int i = 0; i < V; ++i) rdv(est.para_SpeedBias[i], 9);
  for (int i = 1; i < V; ++i) est.pre_integrations[i] = read_preintegration();
  est.initFactorGraph();
  expect_int("init rank", -1, est.last_init_rank);
  for (int j = 1; j < V; ++j) cmp48("vioRelativePoseEdges", -1, est.vioRelativePoseEdges[j]->delta_t, est.vioRelativePoseEdges[j]->delta_R,
                                     est.vioRelativePoseEdges[j]->sqrt_info);
  cmp48("vioPosePriorEdge", -1, est.vioPosePriorEdge->t, est.vioPosePriorEdge->R, est.vioPosePriorEdge->sqrt_info);
  cmp("vioVBPrior.VB", -1, est.vioVBPrior->VB, 9);
  cmp("vioVBPrior.sqrt_info", -1, est.vioVBPrior->sqrt_info, 81);

  // ---- R frames: problemSolve's effect on the factor members comes from the fixture, then
  //      MargForward(); MargBackward(); slideWindow() as backendOptimization does ------------------------
  for (int r = 0; r < R; ++r) {
    const int L = (int)rd();
    rdv(est.para_Pose[0], 7); rdv(est.para_Pose[1], 7); rdv(est.para_Ex_Pose[0], 7);
    for (int k = 0; k < L; ++k) { est.para_Feature[k][0] = rd(); est.MargPointIdx.push_back(k); }
    for (int k = 0; k < L; ++k) {
      auto* pf = new ProjectionFactor();
      rdv(pf->pts_i, 3);
      pf->setIndex(0, 1, k);
      est.forwardProjectiontoSparsify.push_back(pf);
    }
    for (int k = 0; k < L; ++k) rdv(est.forwardProjectiontoSparsify[k]->pts_j, 3);
    // factor members after factor->update(...) and double2vector (the solve itself is out of scope)
    rdv(est.vioPosePriorEdge->t, 3); rdv(est.vioPosePriorEdge->R, 9); rdv(est.vioPosePriorEdge->sqrt_info, 36);
    rdv(est.vioRelativePoseEdges[1]->delta_t, 3); rdv(est.vioRelativePoseEdges[1]->delta_R, 9);
    rdv(est.vioRelativePoseEdges[1]->sqrt_info, 36);
    const int rp_valid = (int)rd();
    const bool have = !est.vioRollPitchEdges.empty() && est.vioRollPitchEdges[0]->index == 0;
    if (have != (rp_valid != 0)) { ++g_fail; fprintf(stderr, "MISMATCH round %d roll-pitch edge bookkeeping\n", r); }
    rdv(est.para_Pose[V - 1], 7); rdv(est.para_SpeedBias[V - 1], 9); rdv(est.para_Pose[V], 7); rdv(est.para_SpeedBias[V], 9);
    rdv(est.vioVBPrior->VB, 9); rdv(est.vioVBPrior->sqrt_info, 81);
    est.Headers[0] = rd();
    rdv(est.Rs[0], 9); rdv(est.Ps[0], 3);
    IntegrationBase* pre = read_preintegration();
    IMUFactor imu(pre);
    imu.setIndex(V - 1, V);
    est.backwardIMUtoSparsify = &imu;

    if (r & 1) {
      est.MargForwardBackward();   // the fused single-event entry point (isv_marg_event): bit-identical results
    } else {
      est.MargForward();
      est.MargBackward();
    }

    expect_int("fwd rank", r, est.last_fwd_rank);
    expect_int("bwd rank", r, est.last_bwd_rank);
    if (est.last_status) { ++g_fail; fprintf(stderr, "round %d status 0x%x\n", r, est.last_status); }
    cmp48("forwardPosePriorEdgeToAdd", r, est.forwardPosePriorEdgeToAdd->t, est.forwardPosePriorEdgeToAdd->R,
          est.forwardPosePriorEdgeToAdd->sqrt_info);
    CombinedFactors* c = est.pose_graph_factors_buf.back();
    cmp48("CombinedFactors.relativePoseFactor", r, c->relativePoseFactor->delta_t, c->relativePoseFactor->delta_R,
          c->relativePoseFactor->sqrt_info);
    cmp("CombinedFactors.covRel", r, c->covRel, 36);
    cmp("CombinedFactors.distance", r, &c->distance, 1);
    cmp("CombinedFactors.covAbs", r, c->covAbs, 4);
    if (c->vio_index != r || c->ts != est.Headers[0]) { ++g_fail; fprintf(stderr, "round %d CombinedFactors bookkeeping\n", r); }
    cmp48("backwardRelativePoseEdgeToAdd", r, est.backwardRelativePoseEdgeToAdd->delta_t, est.backwardRelativePoseEdgeToAdd->delta_R,
          est.backwardRelativePoseEdgeToAdd->sqrt_info);
    cmp("backwardVBEdgeToAdd.VB", r, est.backwardVBEdgeToAdd->VB, 9);
    cmp("backwardVBEdgeToAdd.sqrt_info", r, est.backwardVBEdgeToAdd->sqrt_info, 81);
    cmp("rollPitchFactor.R", r, est.vioRollPitchEdges.back()->R, 9);
    cmp("rollPitchFactor.sqrt_info", r, est.vioRollPitchEdges.back()->sqrt_info, 4);

    est.slideWindowFactors();
    est.backwardIMUtoSparsify = nullptr;
    delete pre;
    // bookkeeping after the rotation: indices of every live edge
    for (int j = 1; j < V; ++j)
      if (est.vioRelativePoseEdges[j]->imu_i != j - 1 || est.vioRelativePoseEdges[j]->imu_j != j) {
        ++g_fail;
        fprintf(stderr, "round %d rel edge %d has indices (%d,%d)\n", r, j, est.vioRelativePoseEdges[j]->imu_i, est.vioRelativePoseEdges[j]->imu_j);
      }
    const int n_rp = (int)rd();
    if ((int)est.vioRollPitchEdges.size() != n_rp) { ++g_fail; fprintf(stderr, "round %d: %zu roll-pitch edges, expected %d\n", r, est.vioRollPitchEdges.size(), n_rp); }
    for (int k = 0; k < n_rp; ++k) expect_int("roll-pitch index", r, k < (int)est.vioRollPitchEdges.size() ? est.vioRollPitchEdges[k]->index : -99);
    if (est.vioPosePriorEdge->index != 0 || est.vioVBPrior->index != V - 1) { ++g_fail; fprintf(stderr, "round %d prior indices\n", r); }
  }
  printf("host_shim_test: %d frames, worst relative error %.3e, %d mismatches, %lld kernel launches\n", R, g_worst, g_fail,
         (long long)isv_launch_count(est.handle()));
  return g_fail ? 1 : 0;
}
