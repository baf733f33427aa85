// C++ `MarginalizationInfo` with VINS-Mono's API -- addResidualBlockInfo / preMarginalize / marginalize /
// getParameterBlocks -- the class BASELINE.json's north_star asks to keep.  IS-VINS deleted it
// (/root/reference has no marginalization_factor.{h,cpp}), so the semantics follow
// VINS-Mono's published vins_estimator/src/factor/marginalization_factor.cpp; parity is unpinned by the
// reference.  Everything numerical runs on the GPU through ONE C-ABI call (isv_marginalize_host:
// Evaluate of every residual block, normal equations, Schur complement, eigen-decomposition); this
// class only keeps the bookkeeping VINS-Mono keeps on the host: block sizes, block order, drop sets.
//
// Differences from VINS-Mono, both forced by the C ABI being factor-type aware instead of calling
// virtual Evaluate():  (1) a ResidualBlockInfo names its factor kind; the factor object is one of the
// classes of isv_estimator_host.hpp;  (2) parameter blocks are ordered by first appearance (VINS-Mono
// iterates an unordered_map keyed by address, i.e. implementation-defined): marginalized blocks of
// local size > 1, marginalized scalars (inverse depths), kept blocks.  parameter_block_idx reports it.
#pragma once
#include <unordered_map>

#include "isv_estimator_host.hpp"

namespace isv_host {

struct YawFactor {                 // include/factor/yaw_factor.h:11-72
  double yaw_meas[3], sqrt_info[1];
  int index = -1;
};

// VINS-Mono projection_td_factor.h (absent from the reference): the 5th parameter block is
// para_Td; rows are already minus ROW / 2.  All visual factors of one problem are either plain or td factors
// (VINS-Mono's ESTIMATE_TD switch).
struct ProjectionTdFactor {
  double pts_i[3], pts_j[3];
  double velocity_i[2], velocity_j[2];
  double td_i = 0, td_j = 0, row_i = 0, row_j = 0;
};
class MarginalizationInfo;
// VINS-Mono marginalization_factor.h `MarginalizationFactor`: the previous round's prior as a residual block; its
// parameter blocks are what prior->getParameterBlocks(addr_shift) returned.
struct MarginalizationFactor {
  explicit MarginalizationFactor(MarginalizationInfo* info) : marginalization_info(info) {}
  MarginalizationInfo* marginalization_info;
};

enum class FactorKind { Projection, IMU, RelativePose, SE3Prior, Linear9, RollPitch, Yaw, ProjectionTd, Marginalization };

struct ResidualBlockInfo {
  ResidualBlockInfo(FactorKind k, const void* cost, std::vector<double*> blocks, std::vector<int> drop)
      : kind(k), cost_function(cost), parameter_blocks(std::move(blocks)), drop_set(std::move(drop)) {}
  FactorKind kind;
  const void* cost_function;               // ProjectionFactor* / IMUFactor* / RelativePoseFactor* / ... (not owned)
  std::vector<double*> parameter_blocks;   // in ceres argument order of the factor
  std::vector<int> drop_set;               // indices into parameter_blocks to marginalize
};

class MarginalizationInfo {
 public:
  // families of the estimator's parameter arrays: para_Pose / para_SpeedBias / para_Ex_Pose / para_Feature
  enum Family { POSE = 0, SPEED_BIAS = 1, EX_POSE = 2, FEATURE = 3, TD = 4 };
  static constexpr int kFamilies = 5;

  // tr_over_row: TR / ROW of ProjectionTdFactor (rolling-shutter read-out per image row), 0 = global shutter
  explicit MarginalizationInfo(isv_handle* h, double cauchy_a = 1.0, double eps_ = 1e-8, double tr_over_row = 0.0)
      : eps(eps_), h_(h), cauchy_a_(cauchy_a), tr_over_row_(tr_over_row) {}
  ~MarginalizationInfo() {
    for (auto& kv : parameter_block_data) delete[] kv.second;
    for (auto* f : factors) delete f;
  }
  MarginalizationInfo(const MarginalizationInfo&) = delete;
  MarginalizationInfo& operator=(const MarginalizationInfo&) = delete;

  static int globalSize(Family f) { return f == POSE || f == EX_POSE ? 7 : (f == SPEED_BIAS ? 9 : 1); }   // FEATURE, TD: 1
  static int localSize(int size) { return size == 7 ? 6 : size; }      // MarginalizationInfo::localSize

  // ceres' SetParameterBlockConstant: the block gets no column
  void setParameterBlockConstant(double* block) { constant_[reinterpret_cast<long>(block)] = true; }

  void addResidualBlockInfo(ResidualBlockInfo* info) {
    // every check runs BEFORE a member is touched: a rejected factor leaves no half-registered state
    if (!info) throw std::runtime_error("addResidualBlockInfo: null ResidualBlockInfo");
    std::vector<Family> fam = families(info->kind);
    if (info->kind == FactorKind::Marginalization) {   // the prior's kept blocks, family by global size of the kept block
      const auto* mf = static_cast<const MarginalizationFactor*>(info->cost_function);
      const MarginalizationInfo* pr = mf->marginalization_info;
      if (prior_factor_) throw std::runtime_error("one MarginalizationFactor per problem");
      if (pr->keep_block_size.size() != info->parameter_blocks.size())
        throw std::runtime_error("MarginalizationFactor: parameter blocks must be prior->getParameterBlocks(addr_shift)");
      fam = pr->keep_block_family_;
    }
    if (fam.size() != info->parameter_blocks.size()) throw std::runtime_error("wrong number of parameter blocks for this factor kind");
    if ((info->kind == FactorKind::ProjectionTd && has_plain_) || (info->kind == FactorKind::Projection && has_td_))
      throw std::runtime_error("ProjectionFactor and ProjectionTdFactor cannot share a problem");
    for (size_t i = 0; i < fam.size(); ++i) {   // a block seen before must come back with the same family (same global size)
      const long addr = reinterpret_cast<long>(info->parameter_blocks[i]);
      auto it = family_.find(addr);
      if (it != family_.end() && it->second != fam[i])
        throw std::runtime_error("a parameter block address is used with two different block families");
    }
    for (int i : info->drop_set) {
      if (i < 0 || (size_t)i >= info->parameter_blocks.size()) throw std::runtime_error("drop_set index out of range");
      if (fam[i] == TD) throw std::runtime_error("para_Td couples with every visual factor: it cannot be marginalized");
      // a scalar block that the previous prior kept cannot be dropped: the prior couples it with the other kept
      // blocks, and scalar blocks are eliminated as a DIAGONAL block (isv_marg_generic.cuh)
      if (info->kind == FactorKind::Marginalization && globalSize(fam[i]) == 1)
        throw std::runtime_error("a scalar block kept by the previous prior cannot be marginalized (diagonal elimination)");
    }
    if (info->kind == FactorKind::Marginalization) prior_factor_ = info;
    factors.push_back(info);
    if (info->kind == FactorKind::ProjectionTd) has_td_ = true;
    if (info->kind == FactorKind::Projection) has_plain_ = true;
    for (size_t i = 0; i < fam.size(); ++i) {
      const long addr = reinterpret_cast<long>(info->parameter_blocks[i]);
      parameter_block_size[addr] = globalSize(fam[i]);
      if (!family_.count(addr)) { family_[addr] = fam[i]; order_.push_back(addr); }
    }
    for (int i : info->drop_set) {
      const long addr = reinterpret_cast<long>(info->parameter_blocks[i]);
      if (!dropped_.count(addr)) { dropped_[addr] = true; drop_order_.push_back(addr); }
    }
  }

  // copies the current value of every parameter block (VINS-Mono also evaluates the factors here; the
  // evaluation happens in marginalize() on the GPU, in the same call as everything else)
  void preMarginalize() {
    for (long addr : order_) {
      if (parameter_block_data.count(addr)) continue;
      const int sz = parameter_block_size[addr];
      double* data = new double[sz];
      std::memcpy(data, reinterpret_cast<double*>(addr), sizeof(double) * sz);
      parameter_block_data[addr] = data;
    }
  }

  void marginalize() {
    // ---- family arrays and indices ----------------------------------------------------------------------
    std::vector<double> arr[kFamilies];
    std::unordered_map<long, int> index;
    std::vector<long> members[kFamilies];
    for (long addr : order_) {
      const Family f = family_[addr];
      index[addr] = (int)members[f].size();
      members[f].push_back(addr);
      const double* v = parameter_block_data.at(addr);
      arr[f].insert(arr[f].end(), v, v + globalSize(f));
    }
    // ---- block order: dense marginalized, scalar marginalized, kept (first appearance) ---------------------
    int pos = 0, m_dense = 0, m_diag = 0;
    parameter_block_idx.clear();
    auto place = [&](long addr) { parameter_block_idx[addr] = pos; pos += localSize(parameter_block_size[addr]); };
    for (long addr : drop_order_)
      if (!constant_.count(addr) && localSize(parameter_block_size[addr]) > 1) { place(addr); }
    m_dense = pos;
    for (long addr : drop_order_)
      if (!constant_.count(addr) && localSize(parameter_block_size[addr]) == 1) { place(addr); ++m_diag; }
    m = pos;
    for (long addr : order_)
      if (!constant_.count(addr) && !dropped_.count(addr)) place(addr);
    n = pos - m;
    std::vector<int32_t> postab[kFamilies];
    for (int f = 0; f < kFamilies; ++f) {
      postab[f].assign(members[f].size(), -1);
      for (size_t k = 0; k < members[f].size(); ++k) {
        auto it = parameter_block_idx.find(members[f][k]);
        if (it != parameter_block_idx.end()) postab[f][k] = it->second;
      }
    }
    // ---- factor lists in the C ABI's layout -----------------------------------------------------------------
    std::vector<int32_t> pidx, iidx, ridx, sidx, vidx, qidx, yidx, tdidx;
    std::vector<double> pobs, ipre, rrec, srec, vrec, qrec, yrec, tdobs;
    size_t P = 0;
    for (auto* f : factors) P += f->kind == FactorKind::Projection || f->kind == FactorKind::ProjectionTd;
    pidx.resize(4 * P); pobs.resize(5 * P);
    if (has_td_) { tdobs.resize(8 * P); tdidx.resize(P); }
    size_t p = 0;
    auto idx_of = [&](double* b) { return index.at(reinterpret_cast<long>(b)); };
    for (auto* f : factors) {
      const auto& pbk = f->parameter_blocks;
      switch (f->kind) {
        case FactorKind::Projection: {
          const auto* c = static_cast<const ProjectionFactor*>(f->cost_function);
          for (int s = 0; s < 4; ++s) pidx[s * P + p] = idx_of(pbk[s]);
          for (int s = 0; s < 3; ++s) pobs[s * P + p] = c->pts_i[s];
          pobs[3 * P + p] = c->pts_j[0]; pobs[4 * P + p] = c->pts_j[1];
          ++p;
          break;
        }
        case FactorKind::ProjectionTd: {
          const auto* c = static_cast<const ProjectionTdFactor*>(f->cost_function);
          for (int s = 0; s < 4; ++s) pidx[s * P + p] = idx_of(pbk[s]);
          tdidx[p] = idx_of(pbk[4]);
          for (int s = 0; s < 3; ++s) pobs[s * P + p] = c->pts_i[s];
          pobs[3 * P + p] = c->pts_j[0]; pobs[4 * P + p] = c->pts_j[1];
          tdobs[0 * P + p] = c->velocity_i[0]; tdobs[1 * P + p] = c->velocity_i[1];
          tdobs[2 * P + p] = c->velocity_j[0]; tdobs[3 * P + p] = c->velocity_j[1];
          tdobs[4 * P + p] = c->td_i; tdobs[5 * P + p] = c->td_j; tdobs[6 * P + p] = c->row_i; tdobs[7 * P + p] = c->row_j;
          ++p;
          break;
        }
        case FactorKind::Marginalization: break;   // handled below
        case FactorKind::IMU: {
          const auto* c = static_cast<const IMUFactor*>(f->cost_function);
          if (c->pre_integration->dirty) throw std::runtime_error("pre-integration record not up to date (Estimator::preintegrated)");
          iidx.push_back(idx_of(pbk[0])); iidx.push_back(idx_of(pbk[2]));
          if (idx_of(pbk[1]) != iidx[iidx.size() - 2] || idx_of(pbk[3]) != iidx.back())
            throw std::runtime_error("IMU factor: pose and speed-bias of a frame must share their index");
          ipre.insert(ipre.end(), c->pre_integration->record, c->pre_integration->record + ISV_PREINT_REC);
          break;
        }
        case FactorKind::RelativePose: {
          const auto* c = static_cast<const RelativePoseFactor*>(f->cost_function);
          ridx.push_back(idx_of(pbk[0])); ridx.push_back(idx_of(pbk[1]));
          push48(rrec, c->delta_t, c->delta_R, c->sqrt_info);
          break;
        }
        case FactorKind::SE3Prior: {
          const auto* c = static_cast<const SE3PriorFactor*>(f->cost_function);
          sidx.push_back(idx_of(pbk[0]));
          push48(srec, c->t, c->R, c->sqrt_info);
          break;
        }
        case FactorKind::Linear9: {
          const auto* c = static_cast<const Linear9Factor*>(f->cost_function);
          vidx.push_back(idx_of(pbk[0]));
          vrec.insert(vrec.end(), c->VB, c->VB + 9); vrec.insert(vrec.end(), c->sqrt_info, c->sqrt_info + 81);
          break;
        }
        case FactorKind::RollPitch: {
          const auto* c = static_cast<const RollPitchFactor*>(f->cost_function);
          qidx.push_back(idx_of(pbk[0]));
          qrec.insert(qrec.end(), c->R, c->R + 9); qrec.insert(qrec.end(), c->sqrt_info, c->sqrt_info + 4);
          break;
        }
        case FactorKind::Yaw: {
          const auto* c = static_cast<const YawFactor*>(f->cost_function);
          yidx.push_back(idx_of(pbk[0]));
          yrec.insert(yrec.end(), c->yaw_meas, c->yaw_meas + 3); yrec.push_back(c->sqrt_info[0]);
          break;
        }
      }
    }
    isv_marg_host_in in;
    std::memset(&in, 0, sizeof(in));
    in.pb = isv_param_blocks{(int32_t)members[POSE].size(), (int32_t)members[SPEED_BIAS].size(), (int32_t)members[EX_POSE].size(),
                             (int32_t)members[FEATURE].size(), arr[POSE].data(), arr[SPEED_BIAS].data(), arr[EX_POSE].data(),
                             arr[FEATURE].data()};
    in.proj.n = (int64_t)P; in.proj.stride = (int64_t)P; in.proj.idx = pidx.data(); in.proj.obs = pobs.data();
    in.proj.cauchy_a = cauchy_a_;
    if (has_td_) {
      in.proj.td_obs = tdobs.data(); in.proj.td = arr[TD].data(); in.proj.td_idx = tdidx.data();
      in.proj.n_td = (int32_t)members[TD].size(); in.proj.tr_over_row = tr_over_row_;
      in.pos_td = postab[TD].data();
    }
    // MarginalizationFactor: blocks in the prior's keep order, located in THIS problem's tangent vector
    std::vector<isv_prior_block> pblk;
    std::vector<double> px0, px;
    isv_marg_prior prior_in;
    if (prior_factor_) {
      const MarginalizationInfo* pr = static_cast<const MarginalizationFactor*>(prior_factor_->cost_function)->marginalization_info;
      for (size_t k = 0; k < pr->keep_block_size.size(); ++k) {
        const long addr = reinterpret_cast<long>(prior_factor_->parameter_blocks[k]);
        const int sz = pr->keep_block_size[k];
        auto it = parameter_block_idx.find(addr);
        pblk.push_back(isv_prior_block{sz, pr->keep_block_idx[k] - pr->m, (int32_t)px0.size(),
                                       it == parameter_block_idx.end() ? -1 : it->second});
        px0.insert(px0.end(), pr->keep_block_data[k], pr->keep_block_data[k] + sz);
        const double* cur = parameter_block_data.at(addr);
        px.insert(px.end(), cur, cur + sz);
      }
      prior_in = isv_marg_prior{pr->n, (int32_t)pblk.size(), pblk.data(), pr->linearized_jacobians.data(),
                                pr->linearized_residuals.data(), px0.data(), px.data()};
      in.prior = &prior_in;
    }
    in.imu = isv_imu_factors{(int32_t)(iidx.size() / 2), iidx.data(), ipre.data()};
    in.small_factors.n_rel = (int32_t)(ridx.size() / 2); in.small_factors.rel_idx = ridx.data(); in.small_factors.rel_rec = rrec.data();
    in.small_factors.n_se3 = (int32_t)sidx.size(); in.small_factors.se3_idx = sidx.data(); in.small_factors.se3_rec = srec.data();
    in.small_factors.n_vb = (int32_t)vidx.size(); in.small_factors.vb_idx = vidx.data(); in.small_factors.vb_rec = vrec.data();
    in.small_factors.n_rp = (int32_t)qidx.size(); in.small_factors.rp_idx = qidx.data(); in.small_factors.rp_rec = qrec.data();
    in.small_factors.n_yaw = (int32_t)yidx.size(); in.small_factors.yaw_idx = yidx.data(); in.small_factors.yaw_rec = yrec.data();
    in.small_factors.cauchy_a = cauchy_a_;
    in.pos_pose = postab[POSE].data(); in.pos_speed_bias = postab[SPEED_BIAS].data();
    in.pos_ex_pose = postab[EX_POSE].data(); in.pos_feature = postab[FEATURE].data();
    in.pos = pos; in.m_dense = m_dense; in.m_diag = m_diag; in.schur_only = 0; in.eps = eps;
    A_red.assign((size_t)n * n, 0.0); b_red.assign(n, 0.0);
    linearized_jacobians.assign((size_t)n * n, 0.0); linearized_residuals.assign(n, 0.0);
    isv_marg_host_out out{A_red.data(), b_red.data(), linearized_jacobians.data(), linearized_residuals.data(), 0, 0};
    check(isv_marginalize_host(h_, &in, &out), "isv_marginalize_host");
    rank = out.rank;
    status = out.status;
  }

  // kept blocks in position order; addr_shift maps old block addresses to the ones of the next window
  std::vector<double*> getParameterBlocks(std::unordered_map<long, double*>& addr_shift) {
    std::vector<double*> keep_block_addr;
    keep_block_size.clear(); keep_block_idx.clear(); keep_block_data.clear(); keep_block_family_.clear();
    std::vector<std::pair<int, long>> kept;
    for (const auto& kv : parameter_block_idx)
      if (kv.second >= m) kept.emplace_back(kv.second, kv.first);
    std::sort(kept.begin(), kept.end());
    for (const auto& pr : kept) {
      keep_block_size.push_back(parameter_block_size[pr.second]);
      keep_block_idx.push_back(pr.first);
      keep_block_data.push_back(parameter_block_data[pr.second]);
      keep_block_family_.push_back(family_[pr.second]);
      keep_block_addr.push_back(addr_shift[pr.second]);
    }
    return keep_block_addr;
  }

  std::vector<ResidualBlockInfo*> factors;
  int m = 0, n = 0, rank = 0, status = 0;
  std::unordered_map<long, int> parameter_block_size;   // global size
  std::unordered_map<long, int> parameter_block_idx;    // local position
  std::unordered_map<long, double*> parameter_block_data;
  std::vector<int> keep_block_size, keep_block_idx;
  std::vector<double*> keep_block_data;
  std::vector<double> linearized_jacobians, linearized_residuals;   // n x n column-major, n
  std::vector<double> A_red, b_red;                                 // the reduced system (diagnostic)
  const double eps;

 private:
  static std::vector<Family> families(FactorKind k) {
    switch (k) {
      case FactorKind::Projection: return {POSE, POSE, EX_POSE, FEATURE};
      case FactorKind::ProjectionTd: return {POSE, POSE, EX_POSE, FEATURE, TD};
      case FactorKind::Marginalization: return {};
      case FactorKind::IMU: return {POSE, SPEED_BIAS, POSE, SPEED_BIAS};
      case FactorKind::RelativePose: return {POSE, POSE};
      case FactorKind::Linear9: return {SPEED_BIAS};
      default: return {POSE};
    }
  }
  static void push48(std::vector<double>& v, const double* t, const double* R, const double* s) {
    v.insert(v.end(), t, t + 3); v.insert(v.end(), R, R + 9); v.insert(v.end(), s, s + 36);
  }
  isv_handle* h_;
  double cauchy_a_, tr_over_row_;
  bool has_td_ = false, has_plain_ = false;
  ResidualBlockInfo* prior_factor_ = nullptr;
  std::vector<Family> keep_block_family_;   // of the kept blocks, filled by getParameterBlocks
  std::unordered_map<long, Family> family_;
  std::unordered_map<long, bool> dropped_, constant_;
  std::vector<long> order_, drop_order_;
};

}  // namespace isv_host
