"""Prints per-factor relative errors of the CUDA path vs the oracle (debug aid; run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from is_vins_b200 import DeviceBatch, MargBackend, capi, pack_events
from oracle import sim
from tests.helpers import compare_event, expected_ranks

be = MargBackend(0)
for L in [150, 33, 1000]:
    ch = sim.make_chain(sim.seed_for(1, 0), L=L, rounds=3)
    batch = pack_events(ch.events)
    db = DeviceBatch(batch, "cuda:0")
    be.marg_window_batch(db, capi.RUN_BOTH)
    be.synchronize()
    out = db.outputs()
    for w, ev in enumerate(ch.events):
        errs = compare_event(out, w, ev)
        print("L", L, "w", w, "rank", out.rank[w], "exp", expected_ranks(ev), "status", out.status[w])
        for k, v in errs.items():
            print("   %-16s %.3e" % (k, v))
        if max(errs.values()) > 1e-6:
            np.set_printoptions(linewidth=200, precision=5)
            print("se3 gpu\n", out.se3_sqrt_info(w), "\nref\n", ev.fwd_out.se3_sqrt_info)
            print("pg gpu\n", out.pg_sqrt_info(w), "\nref\n", ev.fwd_out.pg_sqrt_info)
            print("rel gpu\n", out.rel_sqrt_info(w), "\nref\n", ev.bwd_out.rel_sqrt_info)
            print("vb gpu\n", out.vb_sqrt_info(w), "\nref\n", ev.bwd_out.vb_sqrt_info)
            print("rp gpu\n", out.rp_sqrt_info(w), "\nref\n", ev.bwd_out.rp_sqrt_info)
# quick timing
import torch
ch = sim.make_chain(sim.seed_for(1, 0), L=150, rounds=4)
batch = pack_events(ch.events).tile(1024)
db = DeviceBatch(batch, "cuda:0")
for which, name in [(1, "fwd"), (2, "bwd"), (3, "both")]:
    for _ in range(3):
        be.marg_window_batch(db, which)
    be.synchronize()
    t = time.time()
    for _ in range(10):
        be.marg_window_batch(db, which)
    be.synchronize()
    dt = (time.time() - t) / 10
    print("L=150 n=%d %s: %.3f ms/batch  %.0f windows/s" % (batch.n, name, dt * 1e3, batch.n / dt))
