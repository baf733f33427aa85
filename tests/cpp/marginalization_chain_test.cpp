// Two consecutive MARGIN_OLD rounds through isv_host::MarginalizationInfo, the way VINS-Mono's
// Estimator::optimization() chains them: round 1 marginalizes the oldest frame; getParameterBlocks(addr_shift)
// gives the parameter blocks of the `MarginalizationFactor` that carries the round-1 prior into round 2, where
// the new oldest frame is marginalized.  With use_td every visual factor is a ProjectionTdFactor and para_Td is a
// kept 1-d block (BASELINE configs[3]).  The driver only drives the API and dumps what it got; the numbers are
// checked by tests/test_host_cpp_gpu.py against the oracle's restatement evaluated in the SAME block order.
//   usage: marginalization_chain_test <fixture.bin> <dump.bin>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../is_vins_b200/host/isv_marginalization_info.hpp"

using namespace isv_host;

static std::vector<double> g_data, g_dump;
static size_t g_pos = 0;
static double rd() {
  if (g_pos >= g_data.size()) { fprintf(stderr, "fixture underrun\n"); exit(2); }
  return g_data[g_pos++];
}
static void rdv(double* dst, int n) { for (int i = 0; i < n; ++i) dst[i] = rd(); }
static void dump(double v) { g_dump.push_back(v); }
static void dump(const std::vector<double>& v) { g_dump.insert(g_dump.end(), v.begin(), v.end()); }

struct Visual {
  ProjectionFactor plain;
  ProjectionTdFactor td;
};

static double para_Td[1][1];

// one MARGIN_OLD problem: IMU factor 0 -> 1 and the visual factors hosted in frame 0 (read from the fixture),
// plus whatever prior factors the caller already added
static void add_imu_and_visual(MarginalizationInfo* info, Estimator& est, bool use_td, IntegrationBase& pre, IMUFactor& imu_factor,
                               std::vector<Visual>& vis) {
  rdv(pre.record, ISV_PREINT_REC);
  pre.dirty = false;
  info->addResidualBlockInfo(new ResidualBlockInfo(
      FactorKind::IMU, &imu_factor, {est.para_Pose[0], est.para_SpeedBias[0], est.para_Pose[1], est.para_SpeedBias[1]}, {0, 1}));
  const int P = (int)rd();
  vis.resize(P);
  for (int k = 0; k < P; ++k) {
    const int j = (int)rd(), fidx = (int)rd();
    double pi[3], pj[3];
    rdv(pi, 3); rdv(pj, 3);
    if (use_td) {
      ProjectionTdFactor& f = vis[k].td;
      for (int c = 0; c < 3; ++c) { f.pts_i[c] = pi[c]; f.pts_j[c] = pj[c]; }
      rdv(f.velocity_i, 2); rdv(f.velocity_j, 2);
      f.td_i = rd(); f.td_j = rd(); f.row_i = rd(); f.row_j = rd();
      info->addResidualBlockInfo(new ResidualBlockInfo(
          FactorKind::ProjectionTd, &f, {est.para_Pose[0], est.para_Pose[j], est.para_Ex_Pose[0], est.para_Feature[fidx], para_Td[0]},
          {0, 3}));
    } else {
      ProjectionFactor& f = vis[k].plain;
      for (int c = 0; c < 3; ++c) { f.pts_i[c] = pi[c]; f.pts_j[c] = pj[c]; }
      f.setIndex(0, j, fidx);
      info->addResidualBlockInfo(new ResidualBlockInfo(
          FactorKind::Projection, &f, {est.para_Pose[0], est.para_Pose[j], est.para_Ex_Pose[0], est.para_Feature[fidx]}, {0, 3}));
    }
  }
}

static void dump_round(MarginalizationInfo* info, Estimator& est, int N, int F, const std::vector<double*>& keep) {
  dump(info->m); dump(info->n); dump(info->status); dump(info->rank);
  dump(info->A_red); dump(info->b_red); dump(info->linearized_jacobians); dump(info->linearized_residuals);
  dump((double)keep.size());
  for (size_t k = 0; k < keep.size(); ++k) { dump(info->keep_block_size[k]); dump(info->keep_block_idx[k]); }
  auto idx = [&](double* b) {
    auto it = info->parameter_block_idx.find(reinterpret_cast<long>(b));
    return it == info->parameter_block_idx.end() ? -1.0 : (double)it->second;
  };
  for (int i = 0; i < N; ++i) dump(idx(est.para_Pose[i]));
  for (int i = 0; i < N; ++i) dump(idx(est.para_SpeedBias[i]));
  for (int k = 0; k < F; ++k) dump(idx(est.para_Feature[k]));
  dump(idx(para_Td[0]));
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s fixture.bin dump.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("fixture"); return 2; }
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  g_data.resize(bytes / 8);
  if (fread(g_data.data(), 8, g_data.size(), f) != g_data.size()) return 2;
  fclose(f);

  Estimator est(0);
  const int N = (int)rd(), F = (int)rd();
  const bool use_td = rd() != 0.0;
  const double tr_over_row = rd();
  for (int i = 0; i < N; ++i) rdv(est.para_Pose[i], 7);
  for (int i = 0; i < N; ++i) rdv(est.para_SpeedBias[i], 9);
  rdv(est.para_Ex_Pose[0], 7);
  for (int k = 0; k < F; ++k) est.para_Feature[k][0] = rd();
  para_Td[0][0] = rd();
  double z3[3] = {0, 0, 0};

  // ---- round 1 -----------------------------------------------------------------------------------------------
  MarginalizationInfo* last_marginalization_info = new MarginalizationInfo(est.handle(), 1.0, 1e-8, tr_over_row);
  last_marginalization_info->setParameterBlockConstant(est.para_Ex_Pose[0]);
  IntegrationBase pre1(z3, z3, z3, z3);
  IMUFactor imu1(&pre1);
  std::vector<Visual> vis1;
  add_imu_and_visual(last_marginalization_info, est, use_td, pre1, imu1, vis1);
  SE3PriorFactor se3;
  rdv(se3.t, 3); rdv(se3.R, 9); rdv(se3.sqrt_info, 36);
  last_marginalization_info->addResidualBlockInfo(new ResidualBlockInfo(FactorKind::SE3Prior, &se3, {est.para_Pose[0]}, {0}));
  RelativePoseFactor relf;
  rdv(relf.delta_t, 3); rdv(relf.delta_R, 9); rdv(relf.sqrt_info, 36);
  last_marginalization_info->addResidualBlockInfo(
      new ResidualBlockInfo(FactorKind::RelativePose, &relf, {est.para_Pose[0], est.para_Pose[1]}, {0}));
  last_marginalization_info->preMarginalize();
  last_marginalization_info->marginalize();
  std::unordered_map<long, double*> addr_shift;
  for (int i = 1; i < N; ++i) {   // MARGIN_OLD: block i moves to i - 1
    addr_shift[reinterpret_cast<long>(est.para_Pose[i])] = est.para_Pose[i - 1];
    addr_shift[reinterpret_cast<long>(est.para_SpeedBias[i])] = est.para_SpeedBias[i - 1];
  }
  addr_shift[reinterpret_cast<long>(para_Td[0])] = para_Td[0];
  std::vector<double*> last_marginalization_parameter_blocks = last_marginalization_info->getParameterBlocks(addr_shift);
  dump_round(last_marginalization_info, est, N, F, last_marginalization_parameter_blocks);

  // ---- the window slides and the next solve moves the estimates -------------------------------------------
  for (int i = 0; i + 1 < N; ++i) rdv(est.para_Pose[i], 7);
  for (int i = 0; i + 1 < N; ++i) rdv(est.para_SpeedBias[i], 9);
  for (int k = 0; k < F; ++k) est.para_Feature[k][0] = rd();
  para_Td[0][0] = rd();

  // ---- round 2: the round-1 prior as a MarginalizationFactor -----------------------------------------------
  MarginalizationInfo* marginalization_info = new MarginalizationInfo(est.handle(), 1.0, 1e-8, tr_over_row);
  marginalization_info->setParameterBlockConstant(est.para_Ex_Pose[0]);
  MarginalizationFactor marginalization_factor(last_marginalization_info);
  std::vector<int> drop_set;
  for (int i = 0; i < (int)last_marginalization_parameter_blocks.size(); ++i)
    if (last_marginalization_parameter_blocks[i] == est.para_Pose[0] || last_marginalization_parameter_blocks[i] == est.para_SpeedBias[0])
      drop_set.push_back(i);
  marginalization_info->addResidualBlockInfo(
      new ResidualBlockInfo(FactorKind::Marginalization, &marginalization_factor, last_marginalization_parameter_blocks, drop_set));
  IntegrationBase pre2(z3, z3, z3, z3);
  IMUFactor imu2(&pre2);
  std::vector<Visual> vis2;
  add_imu_and_visual(marginalization_info, est, use_td, pre2, imu2, vis2);
  marginalization_info->preMarginalize();
  marginalization_info->marginalize();
  std::unordered_map<long, double*> addr_shift2;
  for (int i = 1; i + 1 < N; ++i) {
    addr_shift2[reinterpret_cast<long>(est.para_Pose[i])] = est.para_Pose[i - 1];
    addr_shift2[reinterpret_cast<long>(est.para_SpeedBias[i])] = est.para_SpeedBias[i - 1];
  }
  addr_shift2[reinterpret_cast<long>(para_Td[0])] = para_Td[0];
  std::vector<double*> keep2 = marginalization_info->getParameterBlocks(addr_shift2);
  dump_round(marginalization_info, est, N, F, keep2);

  // host latency of one warm marginalize() (pack + H2D + Evaluate + normal equations + Schur + eigen + D2H, blocking)
  double best_us = 1e30;
  for (int it = 0; it < 6; ++it) {
    const auto t0 = std::chrono::steady_clock::now();
    marginalization_info->marginalize();
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    if (it > 0 && us < best_us) best_us = us;
  }
  printf("marginalization_chain_test: warm marginalize() of round 2 (m=%d n=%d): %.0f us\n", marginalization_info->m,
         marginalization_info->n, best_us);

  FILE* o = fopen(argv[2], "wb");
  if (!o || fwrite(g_dump.data(), 8, g_dump.size(), o) != g_dump.size()) { perror("dump"); return 2; }
  fclose(o);
  printf("marginalization_chain_test: round 1 m=%d n=%d rank=%d status=0x%x | round 2 m=%d n=%d rank=%d status=0x%x (%zu + %zu visual factors%s)\n",
         last_marginalization_info->m, last_marginalization_info->n, last_marginalization_info->rank, last_marginalization_info->status,
         marginalization_info->m, marginalization_info->n, marginalization_info->rank, marginalization_info->status, vis1.size(),
         vis2.size(), use_td ? ", ProjectionTdFactor" : "");
  const int bad = last_marginalization_info->status | marginalization_info->status;
  delete marginalization_info;
  delete last_marginalization_info;
  return bad ? 1 : 0;
}
