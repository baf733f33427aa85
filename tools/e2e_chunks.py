#!/usr/bin/env python
"""tools/e2e_chunks.py [L] [n] -- the host-pointer batch call (isv_marg_window_batch_host) alone: ms per call for the three
input ABIs; run under ISV_HOST_CHUNKS=k to sweep the chunk pipeline.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from is_vins_b200 import MargBackend, capi
    from is_vins_b200.backend import xy_as_f32
    from is_vins_b200.batch import WindowOutputs
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
    torch.cuda.set_device(0)
    be = MargBackend(0)
    b = bench.make_batch(L, n, 5)
    keep = []
    for f in b.FIELDS:
        a = getattr(b, f)
        if a is not None:
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            keep.append(t)
            setattr(b, f, t.numpy())
    xyf = torch.from_numpy(xy_as_f32(b.lm_obs)).pin_memory()
    shapes = {"se3": (n, capi.SE3_REC), "pg": (n, capi.PG_REC), "rel": (n, capi.REL_REC), "vb": (n, capi.VB_REC), "rp": (n, capi.RP_REC)}
    ht = {k: torch.zeros(s, dtype=torch.float64).pin_memory() for k, s in shapes.items()}
    ht["rank"] = torch.zeros((n, 2), dtype=torch.int32).pin_memory()
    ht["status"] = torch.zeros((n,), dtype=torch.int32).pin_memory()
    hout = WindowOutputs(*[ht[k].numpy() for k in ("se3", "pg", "rel", "vb", "rp", "rank", "status")])
    res = {"L": L, "n": n, "chunks": os.environ.get("ISV_HOST_CHUNKS", "default")}
    from is_vins_b200.batch import pack_tri_inputs, packed_outputs
    tri_t = {k: (None if v is None else torch.from_numpy(v).pin_memory()) for k, v in pack_tri_inputs(b).items()}
    tri = {k: (None if v is None else v.numpy()) for k, v in tri_t.items()}
    pt = {k: torch.from_numpy(getattr(packed_outputs(n), k)).pin_memory() for k in ("se3", "pg", "rel", "vb", "rp", "rank", "status")}
    pout = WindowOutputs(*[pt[k].numpy() for k in ("se3", "pg", "rel", "vb", "rp", "rank", "status")])
    for tag, kw in (("abi4", dict(raw_imu=True, z_one=True, xy_f32=xyf.numpy(), tri_in=tri, tri_out=True)),
                    ("abi3", dict(raw_imu=True, z_one=True, xy_f32=xyf.numpy())), ("abi2", dict(raw_imu=True, z_one=True)), ("abi1", dict())):
        ho = pout if tag == "abi4" else hout
        for _ in range(3):
            be.marg_window_batch_host(b, capi.RUN_BOTH, ho, **kw)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            be.marg_window_batch_host(b, capi.RUN_BOTH, ho, **kw)
            ts.append(time.perf_counter() - t0)
        res[tag + "_ms"] = round(float(np.median(ts)) * 1e3, 3)
    print(json.dumps(res))
    be.close()


if __name__ == "__main__":
    main()
