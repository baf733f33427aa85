// Drives isv_host::MarginalizationInfo (is_vins_b200/host/isv_marginalization_info.hpp) the way VINS-Mono's
// Estimator::optimization() drives its marginalization_info for MARGIN_OLD: the IMU factor 0 -> 1, every
// projection factor hosted in frame 0 and the prior factors on frame 0 are added with their drop sets, then
// preMarginalize(); marginalize(); getParameterBlocks().  The fixture (tests/test_host_cpp_gpu.py) carries
// the 80-bit extended-precision reduced system of the same problem as the yardstick.
//   usage: marginalization_info_test <fixture.bin>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../is_vins_b200/host/isv_marginalization_info.hpp"

using namespace isv_host;

static std::vector<double> g_data;
static size_t g_pos = 0;
static double rd() {
  if (g_pos >= g_data.size()) { fprintf(stderr, "fixture underrun\n"); exit(2); }
  return g_data[g_pos++];
}
static void rdv(double* dst, int n) { for (int i = 0; i < n; ++i) dst[i] = rd(); }
static double rel(const std::vector<double>& a, const std::vector<double>& b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); ++i) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
  return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s fixture.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("fixture"); return 2; }
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  g_data.resize(bytes / 8);
  if (fread(g_data.data(), 8, g_data.size(), f) != g_data.size()) return 2;
  fclose(f);

  Estimator est(0);
  const int N = (int)rd(), F = (int)rd(), P = (int)rd();
  for (int i = 0; i < N; ++i) rdv(est.para_Pose[i], 7);
  for (int i = 0; i < N; ++i) rdv(est.para_SpeedBias[i], 9);
  rdv(est.para_Ex_Pose[0], 7);
  for (int k = 0; k < F; ++k) est.para_Feature[k][0] = rd();

  MarginalizationInfo* marginalization_info = new MarginalizationInfo(est.handle(), /*cauchy_a=*/1.0);
  marginalization_info->setParameterBlockConstant(est.para_Ex_Pose[0]);   // !ESTIMATE_EXTRINSIC (estimator.cpp:1037)
  // IMU factor 0 -> 1 (its pre-integration record comes with the fixture)
  double z3[3] = {0, 0, 0};
  IntegrationBase pre(z3, z3, z3, z3);
  rdv(pre.record, ISV_PREINT_REC);
  pre.dirty = false;
  IMUFactor imu_factor(&pre);
  marginalization_info->addResidualBlockInfo(new ResidualBlockInfo(
      FactorKind::IMU, &imu_factor, {est.para_Pose[0], est.para_SpeedBias[0], est.para_Pose[1], est.para_SpeedBias[1]}, {0, 1}));
  // projection factors hosted in frame 0
  std::vector<ProjectionFactor> proj(P);
  for (int k = 0; k < P; ++k) {
    const int j = (int)rd(), fidx = (int)rd();
    rdv(proj[k].pts_i, 3); rdv(proj[k].pts_j, 3);
    proj[k].setIndex(0, j, fidx);
    marginalization_info->addResidualBlockInfo(new ResidualBlockInfo(
        FactorKind::Projection, &proj[k], {est.para_Pose[0], est.para_Pose[j], est.para_Ex_Pose[0], est.para_Feature[fidx]}, {0, 3}));
  }
  SE3PriorFactor se3;
  rdv(se3.t, 3); rdv(se3.R, 9); rdv(se3.sqrt_info, 36);
  marginalization_info->addResidualBlockInfo(new ResidualBlockInfo(FactorKind::SE3Prior, &se3, {est.para_Pose[0]}, {0}));
  RelativePoseFactor relf;
  rdv(relf.delta_t, 3); rdv(relf.delta_R, 9); rdv(relf.sqrt_info, 36);
  marginalization_info->addResidualBlockInfo(
      new ResidualBlockInfo(FactorKind::RelativePose, &relf, {est.para_Pose[0], est.para_Pose[1]}, {0}));

  marginalization_info->preMarginalize();
  marginalization_info->marginalize();

  int fail = 0;
  const int m = (int)rd(), n = (int)rd();
  const double tol = rd();
  if (marginalization_info->m != m || marginalization_info->n != n) {
    fprintf(stderr, "MISMATCH m,n = %d,%d expected %d,%d\n", marginalization_info->m, marginalization_info->n, m, n);
    return 1;
  }
  if (marginalization_info->status) { ++fail; fprintf(stderr, "status 0x%x\n", marginalization_info->status); }
  std::vector<double> A_ref((size_t)n * n), b_ref(n);
  rdv(A_ref.data(), n * n);   // column-major
  rdv(b_ref.data(), n);
  const double eA = rel(marginalization_info->A_red, A_ref), eb = rel(marginalization_info->b_red, b_ref);
  if (!(eA <= tol)) { ++fail; fprintf(stderr, "MISMATCH A_red rel err %.3e > %.3e\n", eA, tol); }
  if (!(eb <= tol)) { ++fail; fprintf(stderr, "MISMATCH b_red rel err %.3e > %.3e\n", eb, tol); }
  // the prior it defines: J^T J == A_red (up to the dropped eigenvalues <= eps), J^T r == b_red projected
  const std::vector<double>& J = marginalization_info->linearized_jacobians;
  std::vector<double> JtJ((size_t)n * n, 0.0);
  for (int a = 0; a < n; ++a)
    for (int c = 0; c < n; ++c) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += J[k + (size_t)n * a] * J[k + (size_t)n * c];
      JtJ[a + (size_t)n * c] = s;
    }
  const double eJ = rel(JtJ, marginalization_info->A_red);
  if (!(eJ <= 1e-7)) { ++fail; fprintf(stderr, "MISMATCH J^T J vs A_red %.3e\n", eJ); }
  // parameter block bookkeeping
  if (marginalization_info->parameter_block_idx[reinterpret_cast<long>(est.para_Pose[0])] != 0 ||
      marginalization_info->parameter_block_idx[reinterpret_cast<long>(est.para_SpeedBias[0])] != 6) { ++fail; fprintf(stderr, "dense marginalized blocks misplaced\n"); }
  if (marginalization_info->parameter_block_idx.count(reinterpret_cast<long>(est.para_Ex_Pose[0]))) { ++fail; fprintf(stderr, "constant block got a column\n"); }
  std::unordered_map<long, double*> addr_shift;
  for (int i = 1; i < N; ++i) {   // MARGIN_OLD: block i moves to i - 1 (VINS-Mono Estimator::optimization)
    addr_shift[reinterpret_cast<long>(est.para_Pose[i])] = est.para_Pose[i - 1];
    addr_shift[reinterpret_cast<long>(est.para_SpeedBias[i])] = est.para_SpeedBias[i - 1];
  }
  std::vector<double*> keep = marginalization_info->getParameterBlocks(addr_shift);
  int cols = 0;
  for (size_t k = 0; k < keep.size(); ++k) {
    if (marginalization_info->keep_block_idx[k] != m + cols) { ++fail; fprintf(stderr, "kept block %zu at %d, expected %d\n", k, marginalization_info->keep_block_idx[k], m + cols); }
    cols += MarginalizationInfo::localSize(marginalization_info->keep_block_size[k]);
    if (!keep[k]) { ++fail; fprintf(stderr, "kept block %zu has no shifted address\n", k); }
  }
  if (cols != n || keep.empty() || keep[0] != est.para_Pose[0]) { ++fail; fprintf(stderr, "kept blocks cover %d of %d columns\n", cols, n); }
  printf("marginalization_info_test: m=%d n=%d rank=%d  A_red err %.2e  b_red err %.2e (tol %.1e)  J^T J err %.2e, %d mismatches\n", m, n,
         marginalization_info->rank, eA, eb, tol, eJ, fail);
  delete marginalization_info;
  return fail ? 1 : 0;
}
