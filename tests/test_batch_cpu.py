"""CPU tests of the host-side batch logic, incl. the world_size-2 window sharding over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from is_vins_b200.batch import WindowBatch
from tests.helpers import load_batch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _batch():
    return load_batch(os.path.join(GOLD, "cfg1_L150_ragged.npz"))[0]


def test_tile_and_slice_roundtrip():
    b = _batch()
    t = b.tile(3)
    assert t.n == 3 * b.n and t.n_landmarks == 3 * b.n_landmarks
    for r in range(3):
        s = t.slice(r * b.n, (r + 1) * b.n)
        assert np.array_equal(s.lm_offset, b.lm_offset) and np.array_equal(s.lm_obs, b.lm_obs)
        assert np.array_equal(s.preint, b.preint) and np.array_equal(s.prior_rp, b.prior_rp)
    s = b.slice(2, 5)  # ragged windows, incl. the empty one (L = 0)
    assert s.n == 3 and s.lm_offset[0] == 0
    assert np.array_equal(np.diff(s.lm_offset), np.diff(b.lm_offset)[2:5])
    a = int(b.lm_offset[2])
    assert np.array_equal(s.lm_obs, b.lm_obs[:, a:a + s.n_landmarks])


def test_alg_bytes_formula():
    import bench
    assert bench.alg_bytes(150) == 16312 and bench.alg_bytes(1000) == 63912 and bench.alg_bytes(0, 2) == 5920
    assert bench.alg_bytes(1000, 1, needed_only=True) == 33992


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = _batch()
    n = b.n
    lo, hi = rank * n // world, (rank + 1) * n // world   # contiguous block partition
    shard = b.slice(lo, hi)
    # the only cross-rank traffic of the path: max-reduce of the timing, sum of counts (untimed)
    t = torch.tensor([float(rank + 1), float(shard.n), float(shard.n_landmarks)], dtype=torch.float64)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    q.put((rank, lo, hi, float(mx[0]), float(sm[1]), float(sm[2])))
    dist.barrier()
    dist.destroy_process_group()


def test_window_sharding_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = _batch()
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == b.n     # shards tile [0, n)
    assert all(r[3] == 2.0 for r in res)                                      # max over ranks
    assert all(r[4] == b.n and r[5] == b.n_landmarks for r in res)            # nothing lost or duplicated


def test_tri_record_packing_round_trips_and_matches_the_header():
    """ABI 4 (ISV_IN_TRI_RECORDS / ISV_OUT_TRI_RECORDS): the Python reference packing -- record sizes equal the header's,
    pack -> unpack is the identity on records whose sqrt_info blocks are upper triangular (covRel symmetric), and a non-zero
    strict lower triangle is refused."""
    import os
    import re
    from is_vins_b200 import capi
    from is_vins_b200.batch import TRI_LAYOUTS, _tri_index, pack_tri, unpack_tri
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "isv_capi.h")).read()
    want = {k: int(v) for k, v in re.findall(r"#define (ISV_\w+_TRI_REC) (\d+)", hdr)}
    assert want == {"ISV_SE3_TRI_REC": capi.SE3_TRI_REC, "ISV_REL_TRI_REC": capi.REL_TRI_REC, "ISV_VB_TRI_REC": capi.VB_TRI_REC,
                    "ISV_RP_IN_TRI_REC": capi.RP_IN_TRI_REC, "ISV_PG_TRI_REC": capi.PG_TRI_REC, "ISV_RP_TRI_REC": capi.RP_TRI_REC}
    sizes = {"se3": (capi.SE3_REC, capi.SE3_TRI_REC), "rel": (capi.REL_REC, capi.REL_TRI_REC), "vb": (capi.VB_REC, capi.VB_TRI_REC),
             "rp_in": (capi.RP_IN_REC, capi.RP_IN_TRI_REC), "pg": (capi.PG_REC, capi.PG_TRI_REC), "rp": (capi.RP_REC, capi.RP_TRI_REC)}
    rng = np.random.default_rng(3)
    for fam, (full, packed) in sizes.items():
        idx, fl = _tri_index(TRI_LAYOUTS[fam])
        assert fl == full and len(idx) == packed and len(set(idx.tolist())) == packed
        rec = np.zeros((7, full))
        fo = 0
        for sno, (kind, N) in enumerate(TRI_LAYOUTS[fam]):
            if kind == 0:
                rec[:, fo:fo + N] = rng.normal(size=(7, N))
                fo += N
            else:
                for w in range(7):
                    M = np.triu(rng.normal(size=(N, N)))
                    if fam == "pg" and sno == 2:
                        M = M + M.T                      # covRel: symmetric
                    rec[w, fo:fo + N * N] = M.flatten(order="F")
                fo += N * N
        p = pack_tri(rec, fam)
        assert p.shape == (7, packed)
        back = unpack_tri(p, fam, symmetric_blocks=(2,) if fam == "pg" else ())
        assert np.array_equal(back, rec), fam
    bad = np.zeros((1, capi.SE3_REC))
    bad[0, 12 + 1] = 1.0                                  # element (1, 0) of the column-major sqrt_info: strict lower triangle
    with pytest.raises(ValueError):
        pack_tri(bad, "se3")
