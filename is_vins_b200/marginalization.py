"""`MarginalizationInfo`: the VINS-Mono marginalization API the north_star names
(addResidualBlockInfo / preMarginalize / marginalize / getParameterBlocks), on the GPU engine
(include/isv_capi.h `isv_eval_problem` + `isv_marginalize_generic`).  IS-VINS deleted this class
(SURVEY.md section 0): the semantics follow VINS-Mono's published marginalization_factor.cpp, with one
documented difference -- parameter blocks are ordered by first appearance (VINS-Mono iterates an
`unordered_map` keyed by address, i.e. implementation-defined order): marginalized blocks of size > 1,
then marginalized scalar blocks (inverse depths), then the kept blocks.  `parameter_block_idx` reports
the resulting positions, so the mapping is exact and reproducible.

Parameter blocks are named by (family, index): ("pose", i), ("speed_bias", i), ("ex_pose", e),
("feature", f) -- the `para_Pose[i]` ... arrays of the estimator.  No arithmetic happens on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from .evaluate import DeviceProblem, FactorProblem, eval_problem, rel_record, rp_record, vb_record, yaw_record

Key = Tuple[str, int]
GLOBAL_SIZE = {"pose": 7, "speed_bias": 9, "ex_pose": 7, "feature": 1, "td": 1}
LOCAL_SIZE = {"pose": 6, "speed_bias": 9, "ex_pose": 6, "feature": 1, "td": 1}     # MarginalizationInfo::localSize
# per factor kind: parameter families in ceres argument order, residual size, (offset, row stride) of each
# Jacobian block inside the kind's Evaluate output record
KINDS = {
    "projection": (("pose", "pose", "ex_pose", "feature"), 2, None),
    # VINS-Mono ProjectionTdFactor (5th block: the camera-IMU time offset para_Td); members pts_i, pts_j,
    # velocity_i, velocity_j, td_i, td_j, row_i, row_j (rows already minus ROW / 2)
    "projection_td": (("pose", "pose", "ex_pose", "feature", "td"), 2, None),
    "imu": (("pose", "speed_bias", "pose", "speed_bias"), 15, ((0, 7), (105, 9), (240, 7), (345, 9))),
    "rel": (("pose", "pose"), 6, ((0, 7), (42, 7))),
    "se3": (("pose",), 6, ((0, 7),)),
    "vb": (("speed_bias",), 9, ((0, 9),)),
    "rp": (("pose",), 2, ((0, 7),)),
    "yaw": (("pose",), 1, ((0, 7),)),
    # VINS-Mono MarginalizationFactor: the previous round's prior (member `prior` = the marginalized
    # MarginalizationInfo); parameter blocks = its kept blocks, in getParameterBlocks() order, under the keys
    # they have NOW (after the window slid: VINS-Mono's addr_shift)
    "marginalization": (None, None, None),
}


class ResidualBlockInfo:
    """cost function (by kind + its members), the parameter blocks it touches and the drop set."""

    def __init__(self, kind: str, parameter_blocks: Sequence[Key], drop_set: Sequence[int] = (), **members):
        fam = KINDS[kind][0]
        if kind == "marginalization":
            keep = members["prior"].getParameterBlocks()
            assert [GLOBAL_SIZE[k[0]] for k in parameter_blocks] == [size for _, size, _ in keep], \
                "parameter blocks must be the prior's kept blocks, in getParameterBlocks() order"
        else:
            assert len(parameter_blocks) == len(fam) and all(k[0] == f for k, f in zip(parameter_blocks, fam)), \
                f"{kind} takes parameter blocks {fam}"
        self.kind, self.parameter_blocks, self.drop_set, self.members = kind, list(parameter_blocks), list(drop_set), members


class PriorState:
    """A marginalization prior outside the `MarginalizationInfo` that produced it -- what a checkpoint of
    VINS-Mono's `last_marginalization_info` holds (linearized_jacobians / linearized_residuals / keep_block_size /
    keep_block_idx / keep_block_data).  Usable as the `prior` member of a "marginalization" residual block.
    keys: the kept blocks in position order; x0: their values at the linearization point."""

    def __init__(self, keys: Sequence[Key], linearized_jacobians, linearized_residuals, x0: Sequence[np.ndarray]):
        self.linearized_jacobians = np.ascontiguousarray(linearized_jacobians, float)
        self.linearized_residuals = np.ascontiguousarray(linearized_residuals, float)
        self.n, self.m = int(self.linearized_residuals.shape[0]), 0
        self._keep, pos = [], 0
        for k in keys:
            self._keep.append((k, GLOBAL_SIZE[k[0]], pos))
            pos += LOCAL_SIZE[k[0]]
        assert pos == self.n and self.linearized_jacobians.shape == (self.n, self.n)
        self.keep_block_data = {k: np.asarray(v, float).reshape(-1).copy() for k, v in zip(keys, x0)}

    @classmethod
    def from_info(cls, mi: "MarginalizationInfo") -> "PriorState":
        keep = mi.getParameterBlocks()
        return cls([k for k, _, _ in keep], mi.linearized_jacobians, mi.linearized_residuals,
                   [mi.keep_block_data[k] for k, _, _ in keep])

    def getParameterBlocks(self, addr_shift: Optional[Dict[Key, Key]] = None):
        if addr_shift is not None:
            return [addr_shift[k] for k, _, _ in self._keep]
        return list(self._keep)


class isv_ne_block(C.Structure):
    _fields_ = [("jac_offset", C.c_int64), ("row_stride", C.c_int32), ("local_size", C.c_int32), ("pos", C.c_int32),
                ("reserved", C.c_int32)]


class isv_ne_factor(C.Structure):
    _fields_ = [("res_offset", C.c_int64), ("n_res", C.c_int32), ("n_blocks", C.c_int32), ("first_block", C.c_int32),
                ("problem", C.c_int32)]


class MarginalizationInfo:
    def __init__(self, backend, eps: float = 1e-8, cauchy_a: float = 1.0, constant: Sequence[Key] = (),
                 tr_over_row: float = 0.0):
        """cauchy_a: CauchyLoss scale of every non-IMU residual block (VINS passes `loss_function` to the
        projection and prior blocks and NULL to the IMU block); 0 disables.  constant: parameter blocks
        ceres holds constant (SetParameterBlockConstant) -- they get no column."""
        self.be, self.eps, self.cauchy_a, self.constant = backend, float(eps), float(cauchy_a), set(constant)
        self.tr_over_row = float(tr_over_row)      # TR / ROW of ProjectionTdFactor (rolling shutter), 0 = global
        self.factors: List[ResidualBlockInfo] = []
        self.parameter_block_size: Dict[Key, int] = {}
        self.parameter_block_idx: Dict[Key, int] = {}
        self._drop: List[Key] = []
        self.m = self.n = 0
        self._dp: Optional[DeviceProblem] = None

    # ---- VINS-Mono API ------------------------------------------------------------------------------
    def addResidualBlockInfo(self, info: ResidualBlockInfo) -> None:
        # validate first: a rejected factor leaves no half-registered state
        for i in info.drop_set:
            if not 0 <= i < len(info.parameter_blocks):
                raise ValueError(f"drop_set index {i} out of range for {len(info.parameter_blocks)} parameter blocks")
            if info.parameter_blocks[i][0] == "td":
                raise ValueError("the time offset couples with every ProjectionTdFactor: it cannot be a diagonal "
                                 "marginalized scalar (VINS-Mono never marginalizes para_Td either)")
        self.factors.append(info)
        for k in info.parameter_blocks:
            self.parameter_block_size[k] = GLOBAL_SIZE[k[0]]
        for i in info.drop_set:
            k = info.parameter_blocks[i]
            if k not in self._drop:
                self._drop.append(k)

    def preMarginalize(self, para: Dict[str, np.ndarray]) -> None:
        """Evaluate every residual block at the current parameter values (GPU, isv_eval_problem)."""
        by = {k: [f for f in self.factors if f.kind == k] for k in KINDS}
        self._by = by
        P = len(by["projection"])
        pidx = np.zeros((4, P), np.int32)
        pobs = np.zeros((5, P))
        for c, f in enumerate(by["projection"]):
            pidx[:, c] = [k[1] for k in f.parameter_blocks]
            pobs[0:3, c], pobs[3:5, c] = f.members["pts_i"], np.asarray(f.members["pts_j"])[0:2]
        i32 = lambda rows, shape: np.asarray(rows, np.int32).reshape(shape)
        f64 = lambda rows, w: np.asarray(rows, float).reshape(-1, w)
        M = lambda f: f.members
        fp = FactorProblem(
            np.ascontiguousarray(para["pose"], float), np.ascontiguousarray(para["speed_bias"], float),
            np.ascontiguousarray(para["ex_pose"], float).reshape(-1, 7), np.ascontiguousarray(para["feature"], float),
            pidx, pobs,
            i32([(f.parameter_blocks[0][1], f.parameter_blocks[2][1]) for f in by["imu"]], (-1, 2)),
            f64([M(f)["preint"] for f in by["imu"]], capi.PREINT_REC),
            i32([(f.parameter_blocks[0][1], f.parameter_blocks[1][1]) for f in by["rel"]], (-1, 2)),
            f64([rel_record(M(f)["delta_t"], M(f)["delta_R"], M(f)["sqrt_info"]) for f in by["rel"]], capi.REL_REC),
            i32([f.parameter_blocks[0][1] for f in by["se3"]], (-1,)),
            f64([rel_record(M(f)["t"], M(f)["R"], M(f)["sqrt_info"]) for f in by["se3"]], capi.SE3_REC),
            i32([f.parameter_blocks[0][1] for f in by["vb"]], (-1,)),
            f64([vb_record(M(f)["VB"], M(f)["sqrt_info"]) for f in by["vb"]], capi.VB_REC),
            i32([f.parameter_blocks[0][1] for f in by["rp"]], (-1,)),
            f64([rp_record(M(f)["R"], M(f)["sqrt_info"]) for f in by["rp"]], capi.RP_REC),
            i32([f.parameter_blocks[0][1] for f in by["yaw"]], (-1,)),
            f64([yaw_record(M(f)["yaw_meas"], M(f)["sqrt_info"]) for f in by["yaw"]], capi.YAW_REC))
        for f in by["imu"]:
            assert f.parameter_blocks[0][1] == f.parameter_blocks[1][1] and f.parameter_blocks[2][1] == f.parameter_blocks[3][1]
        self._dp = DeviceProblem(fp, f"cuda:{self.be.device}")
        eval_problem(self.be, self._dp, self.cauchy_a)
        self._para = {k: np.array(v, float, copy=True).reshape((-1, GLOBAL_SIZE[k]) if GLOBAL_SIZE[k] > 1 else (-1,))
                      for k, v in para.items()}
        # MarginalizationFactor (at most one, like VINS-Mono): residual = r0 + J dx on the GPU (isv_eval_marg_prior)
        self._prior = None
        assert len(by["marginalization"]) <= 1, "one MarginalizationFactor per problem"
        for f in by["marginalization"]:
            import torch
            dev = self._dp.device
            prior = f.members["prior"]
            keep = prior.getParameterBlocks()
            n = prior.n
            x0 = np.concatenate([np.asarray(prior.keep_block_data[k], float).reshape(-1) for k, _, _ in keep])
            x = np.concatenate([self._block_value(k) for k in f.parameter_blocks])
            offs = np.concatenate([[0], np.cumsum([size for _, size, _ in keep])])
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a, float)).to(dev)
            d = {"J": t(prior.linearized_jacobians.T.reshape(-1)),        # column-major
                 "r0": t(prior.linearized_residuals), "x0": t(x0), "x": t(x),
                 "res": torch.zeros((n,), dtype=torch.float64, device=dev),
                 "jac": torch.zeros((n * int(offs[-1]),), dtype=torch.float64, device=dev)}
            self._prior = {"f": f, "keep": keep, "offs": offs, "d": d}
            st = self._prior_struct(pos_of=lambda k: -1)
            capi.check(self.be.lib.isv_eval_marg_prior(self.be.h, C.byref(st), C.c_void_p(d["res"].data_ptr()),
                                                       C.c_void_p(d["jac"].data_ptr()),
                                                       C.c_void_p(self._dp.status.data_ptr())), "isv_eval_marg_prior")
        # ProjectionTdFactors: the same kernel with td_obs set (eval_projection_kernel<true>)
        tdf = by["projection_td"]
        self._td = None
        if tdf:
            import torch
            dev = self._dp.device
            T = len(tdf)
            tidx = np.zeros((4, T), np.int32)
            tobs = np.zeros((5, T))
            tdo = np.zeros((8, T))
            tdi = np.zeros((T,), np.int32)
            for c, f in enumerate(tdf):
                tidx[:, c] = [k[1] for k in f.parameter_blocks[:4]]
                tdi[c] = f.parameter_blocks[4][1]
                m_ = f.members
                tobs[0:3, c], tobs[3:5, c] = m_["pts_i"], np.asarray(m_["pts_j"])[0:2]
                tdo[0:2, c], tdo[2:4, c] = np.asarray(m_["velocity_i"])[0:2], np.asarray(m_["velocity_j"])[0:2]
                tdo[4, c], tdo[5, c], tdo[6, c], tdo[7, c] = m_["td_i"], m_["td_j"], m_["row_i"], m_["row_j"]
            t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
            d = {"idx": t(tidx, torch.int32), "obs": t(tobs, torch.float64), "tdo": t(tdo, torch.float64),
                 "tdi": t(tdi, torch.int32), "td": t(np.asarray(para["td"], float).reshape(-1), torch.float64)}
            o = {k: torch.zeros((T, w), dtype=torch.float64, device=dev)
                 for k, w in (("res", 2), ("ji", 14), ("jj", 14), ("je", 14), ("jf", 2), ("jt", 2))}
            pb = self._dp.param_blocks()
            pf = capi.isv_proj_factors(T, T, d["idx"].data_ptr(), d["obs"].data_ptr(), self.cauchy_a, d["tdo"].data_ptr(),
                                       d["td"].data_ptr(), d["tdi"].data_ptr(), int(d["td"].numel()), 0, self.tr_over_row)
            po = capi.isv_proj_eval(*[o[k].data_ptr() for k in ("res", "ji", "jj", "je", "jf", "jt")])
            capi.check(self.be.lib.isv_eval_projection_batch(self.be.h, C.byref(pb), C.byref(pf), C.byref(po),
                                                             C.c_void_p(self._dp.status.data_ptr())),
                       "isv_eval_projection_batch")
            self._td = (d, o)

    def _block_value(self, k: Key) -> np.ndarray:
        return np.asarray(self._para[k[0]][k[1]], float).reshape(-1)

    def _prior_struct(self, pos_of) -> "capi.isv_marg_prior":
        import torch
        pr = self._prior
        blocks = (capi.isv_prior_block * len(pr["keep"]))(*[
            capi.isv_prior_block(size, idx, int(pr["offs"][c]), int(pos_of(key)))
            for c, ((_, size, idx), key) in enumerate(zip(pr["keep"], pr["f"].parameter_blocks))])
        pr["d"]["blocks"] = torch.frombuffer(bytearray(bytes(blocks)), dtype=torch.uint8).to(self._dp.device)
        d = pr["d"]
        return capi.isv_marg_prior(pr["f"].members["prior"].n, len(pr["keep"]), d["blocks"].data_ptr(), d["J"].data_ptr(),
                                   d["r0"].data_ptr(), d["x0"].data_ptr(), d["x"].data_ptr())

    @property
    def prior_residuals(self) -> np.ndarray:
        return self._prior["d"]["res"].cpu().numpy()

    @property
    def prior_jacobians(self) -> List[np.ndarray]:
        """ceres layout: per kept block an (n x global size) row-major Jacobian"""
        pr = self._prior
        flat, n = pr["d"]["jac"].cpu().numpy(), pr["f"].members["prior"].n
        return [flat[n * int(pr["offs"][c]):n * int(pr["offs"][c + 1])].reshape(n, size)
                for c, (_, size, _) in enumerate(pr["keep"])]

    def marginalize(self, keep_tables: bool = False, schur_only: bool = False, build_only: bool = False) -> None:
        """schur_only: stop after the Schur complement (A_red, b_red) -- with every feature dropped and no dense
        block this is the reduced camera system of DENSE_SCHUR (`isv_reduced_system`).  build_only: stop after the
        normal equations (isv_build_normal_equations [+ prior]): `A_dev` [pos, pos], `b_dev` [pos] stay on the device."""
        import torch
        assert self._dp is not None, "call preMarginalize first"
        dp, by = self._dp, self._by
        pos, m_dense, diag = self.order_blocks()
        # ---- one `values` array: the Evaluate outputs, concatenated ---------------------------------------
        names = ["proj_res", "proj_ji", "proj_jj", "proj_je", "proj_jf", "imu_res", "imu_jac", "rel_res", "rel_jac",
                 "se3_res", "se3_jac", "vb_res", "vb_jac", "rp_res", "rp_jac", "yaw_res", "yaw_jac"]
        base, off = {}, 0
        for nm in names:
            base[nm] = off
            off += dp.out[nm].numel()
        parts = [dp.out[nm].reshape(-1) for nm in names]
        if self._td is not None:
            for nm in ("res", "ji", "jj", "je", "jf", "jt"):
                base["td_" + nm] = off
                off += self._td[1][nm].numel()
                parts.append(self._td[1][nm].reshape(-1))
        values = torch.cat(parts)
        facs, blks = [], []
        for c, f in enumerate(by["projection_td"]):
            first = len(blks)
            for nm, k, w in (("td_ji", f.parameter_blocks[0], 14), ("td_jj", f.parameter_blocks[1], 14),
                             ("td_je", f.parameter_blocks[2], 14), ("td_jf", f.parameter_blocks[3], 2),
                             ("td_jt", f.parameter_blocks[4], 2)):
                if k in self.constant:
                    continue
                blks.append((base[nm] + w * c, GLOBAL_SIZE[k[0]], LOCAL_SIZE[k[0]], self.parameter_block_idx[k]))
            facs.append((base["td_res"] + 2 * c, 2, len(blks) - first, first))
        for c, f in enumerate(by["projection"]):
            first = len(blks)
            for nm, k, w in (("proj_ji", f.parameter_blocks[0], 14), ("proj_jj", f.parameter_blocks[1], 14),
                             ("proj_je", f.parameter_blocks[2], 14), ("proj_jf", f.parameter_blocks[3], 2)):
                if k in self.constant:
                    continue
                gs = GLOBAL_SIZE[k[0]]
                blks.append((base[nm] + w * c, gs, LOCAL_SIZE[k[0]], self.parameter_block_idx[k]))
            facs.append((base["proj_res"] + 2 * c, 2, len(blks) - first, first))
        for kind in ("imu", "rel", "se3", "vb", "rp", "yaw"):
            fam, nres, layout = KINDS[kind]
            width = dp.out[kind + "_jac"].shape[1] if dp.out[kind + "_jac"].ndim == 2 else 0
            for c, f in enumerate(by[kind]):
                first = len(blks)
                for k, (o, stride) in zip(f.parameter_blocks, layout):
                    if k in self.constant:
                        continue
                    blks.append((base[kind + "_jac"] + width * c + o, stride, LOCAL_SIZE[k[0]], self.parameter_block_idx[k]))
                facs.append((base[kind + "_res"] + nres * c, nres, len(blks) - first, first))
        fa = (isv_ne_factor * len(facs))(*[isv_ne_factor(r, nr, nb, fb, 0) for r, nr, nb, fb in facs])
        ba = (isv_ne_block * len(blks))(*[isv_ne_block(j, st, ls, p, 0) for j, st, ls, p in blks])
        dev = dp.device
        d_f = torch.frombuffer(bytearray(bytes(fa)), dtype=torch.uint8).to(dev)
        d_b = torch.frombuffer(bytearray(bytes(ba)), dtype=torch.uint8).to(dev)
        n = self.n
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=dev)
        o = {"A": z(pos, pos), "b": z(pos), "A_red": z(n, n), "b_red": z(n), "J": z(n, n), "r": z(n),
             "rank": torch.zeros((1,), dtype=torch.int32, device=dev), "status": torch.zeros((1,), dtype=torch.int32, device=dev)}

        class _In(C.Structure):
            _fields_ = [("n_problems", C.c_int32), ("pos", C.c_int32), ("m_dense", C.c_int32), ("m_diag", C.c_int32),
                        ("n_factors", C.c_int64), ("factors", C.c_void_p), ("blocks", C.c_void_p), ("values", C.c_void_p),
                        ("eps", C.c_double)]

        class _Out(C.Structure):
            _fields_ = [(k, C.c_void_p) for k in ("A", "b", "A_red", "b_red", "linearized_jacobians",
                                                  "linearized_residuals", "rank", "status")]
        gi = _In(1, pos, m_dense, len(diag), len(facs), d_f.data_ptr(), d_b.data_ptr(), values.data_ptr(), self.eps)
        go = _Out(o["A"].data_ptr(), o["b"].data_ptr(), o["A_red"].data_ptr(), o["b_red"].data_ptr(), o["J"].data_ptr(),
                  o["r"].data_ptr(), o["rank"].data_ptr(), o["status"].data_ptr())
        lib = self.be.lib
        if build_only:
            capi.check(lib.isv_build_normal_equations(self.be.h, C.byref(gi), C.byref(go)), "isv_build_normal_equations")
            if self._prior is not None:
                st = self._prior_struct(pos_of=lambda k: -1 if k in self.constant else self.parameter_block_idx[k])
                capi.check(lib.isv_add_marg_prior(self.be.h, C.byref(st), C.c_void_p(self._prior["d"]["res"].data_ptr()),
                                                  C.byref(gi), C.byref(go), 0), "isv_add_marg_prior")
            self.be.synchronize()
            self.pos, self.A_dev, self.b_dev = pos, o["A"], o["b"]
            self.status = int(o["status"].item()) | int(dp.status.item())
            return
        if self._prior is None:
            fn = lib.isv_reduced_system if schur_only else lib.isv_marginalize_generic
            capi.check(fn(self.be.h, C.byref(gi), C.byref(go)), "isv_marginalize_generic")
        else:   # normal equations of the ordinary blocks, + J^T J / J^T r of the prior, then Schur + eigen
            st = self._prior_struct(pos_of=lambda k: -1 if k in self.constant else self.parameter_block_idx[k])
            capi.check(lib.isv_build_normal_equations(self.be.h, C.byref(gi), C.byref(go)), "isv_build_normal_equations")
            capi.check(lib.isv_add_marg_prior(self.be.h, C.byref(st), C.c_void_p(self._prior["d"]["res"].data_ptr()),
                                              C.byref(gi), C.byref(go), 0), "isv_add_marg_prior")
            capi.check(lib.isv_schur_eig(self.be.h, C.byref(gi), C.byref(go), 1 if schur_only else 0), "isv_schur_eig")
        self.be.synchronize()
        if keep_tables:   # for tools/bench_marg_generic.py: the device tables of this problem
            self._gi, self._go = gi, go
            self._tables = {"factors_bytes": bytes(fa), "d_f": d_f, "d_b": d_b, "values": values, "out": o}
        # column-major (Eigen) -> numpy
        self.A_red = o["A_red"].cpu().numpy().reshape(n, n).T.copy()
        self.b_red = o["b_red"].cpu().numpy()
        self.linearized_jacobians = o["J"].cpu().numpy().reshape(n, n).T.copy()
        self.linearized_residuals = o["r"].cpu().numpy()
        self.rank = int(o["rank"].item())
        self.status = int(o["status"].item()) | int(dp.status.item())
        self.pos = pos
        # keep_block_data: the kept blocks at this linearization point (what MarginalizationFactor needs later)
        self.keep_block_data = {k: self._block_value(k).copy() for k, i in self.parameter_block_idx.items() if i >= self.m}

    def order_blocks(self):
        """The block order of the tangent vector (see module docstring): marginalized blocks of local size > 1 in
        drop order, marginalized scalars (the diagonal block), kept blocks by first appearance; blocks held constant
        get no column.  Sets parameter_block_idx, m, n; returns (pos, m_dense, [diagonal keys]).  Pure host logic."""
        order: List[Key] = [k for k in self._drop if LOCAL_SIZE[k[0]] > 1 and k not in self.constant]
        m_dense = sum(LOCAL_SIZE[k[0]] for k in order)
        diag = [k for k in self._drop if LOCAL_SIZE[k[0]] == 1 and k not in self.constant]
        # scalar blocks are eliminated as a DIAGONAL block (only A[d, d] is read): a scalar the previous prior kept is
        # coupled with the prior's other blocks and cannot go there (the kernels raise ISV_W_DIAG_COUPLED for raw callers)
        prior_keys = {k for f in self.factors if f.kind == "marginalization" for k in f.parameter_blocks}
        coupled = [k for k in diag if k in prior_keys]
        if coupled:
            raise ValueError(f"scalar blocks kept by the previous prior cannot be marginalized as a diagonal block: {coupled}")
        order += diag
        seen = set(order)
        for f in self.factors:
            for k in f.parameter_blocks:
                if k not in seen and k not in self.constant:
                    seen.add(k)
                    order.append(k)
        pos = 0
        self.parameter_block_idx = {}
        for k in order:
            self.parameter_block_idx[k] = pos
            pos += LOCAL_SIZE[k[0]]
        self.m, self.n = m_dense + len(diag), pos - m_dense - len(diag)
        return pos, m_dense, diag

    def getParameterBlocks(self, addr_shift: Optional[Dict[Key, Key]] = None):
        """kept blocks in order: (key, global size, position in the reduced tangent vector = idx - m).  With
        addr_shift (VINS-Mono: where each kept block lives after the window slides) returns just the shifted keys:
        the parameter blocks of the MarginalizationFactor of the next problem."""
        keep = [(k, self.parameter_block_size[k], i - self.m) for k, i in self.parameter_block_idx.items() if i >= self.m]
        keep = sorted(keep, key=lambda t: t[2])
        if addr_shift is not None:
            return [addr_shift[k] for k, _, _ in keep]
        return keep


def add_margin_old_blocks(mi: MarginalizationInfo, fp: FactorProblem, td_obs: Optional[np.ndarray] = None,
                          prior=None, prior_keys: Optional[Sequence[Key]] = None) -> List[ResidualBlockInfo]:
    """The residual blocks VINS-Mono's Estimator::optimization() hands to its MarginalizationInfo for MARGIN_OLD
    (published estimator.cpp; not in the reference): the previous prior with the oldest frame's blocks in its drop
    set, the IMU factor 0 -> 1 dropping pose 0 / speed-bias 0, and every visual factor hosted in frame 0 dropping
    the host pose and the feature -- ProjectionTdFactor (5th block para_Td) when `td_obs` [8,P] is given."""
    added = []

    def add(info):
        mi.addResidualBlockInfo(info)
        added.append(info)
    if prior is not None:
        keys = list(prior_keys if prior_keys is not None else [k for k, _, _ in prior.getParameterBlocks()])
        drop = [c for c, k in enumerate(keys) if k in (("pose", 0), ("speed_bias", 0))]
        add(ResidualBlockInfo("marginalization", keys, drop_set=drop, prior=prior))
    add(ResidualBlockInfo("imu", [("pose", 0), ("speed_bias", 0), ("pose", 1), ("speed_bias", 1)], drop_set=[0, 1],
                          preint=fp.imu_preint[0]))
    for k in np.nonzero(fp.proj_idx[0] == 0)[0]:
        i, j, e, f = [int(x) for x in fp.proj_idx[:, k]]
        keys = [("pose", i), ("pose", j), ("ex_pose", e), ("feature", f)]
        pts_i, pts_j = fp.proj_obs[0:3, k].copy(), np.array([fp.proj_obs[3, k], fp.proj_obs[4, k], 1.0])
        if td_obs is None:
            add(ResidualBlockInfo("projection", keys, drop_set=[0, 3], pts_i=pts_i, pts_j=pts_j))
        else:
            t = td_obs[:, k]
            add(ResidualBlockInfo("projection_td", keys + [("td", 0)], drop_set=[0, 3], pts_i=pts_i, pts_j=pts_j,
                                  velocity_i=t[0:2].copy(), velocity_j=t[2:4].copy(), td_i=float(t[4]), td_j=float(t[5]),
                                  row_i=float(t[6]), row_j=float(t[7])))
    return added
