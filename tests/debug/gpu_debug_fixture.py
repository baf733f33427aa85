import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from is_vins_b200 import DeviceBatch, MargBackend, capi
from tests.helpers import load_batch, rel_err
be = MargBackend(0)
for name in sys.argv[1:]:
    batch, ref, _ = load_batch(os.path.join("tests/golden", name))
    db = DeviceBatch(batch, "cuda:0")
    be.marg_window_batch(db, capi.RUN_BOTH); be.synchronize()
    out = db.outputs()
    for w in range(batch.n):
        print(name, w, "L", int(batch.lm_offset[w+1]-batch.lm_offset[w]), "rank", out.rank[w], ref.rank[w], "status", out.status[w],
              "rel %.2e vb %.2e rp %.2e se3 %.2e pg %.2e" % (rel_err(out.rel_sqrt_info(w), ref.rel_sqrt_info(w)), rel_err(out.vb_sqrt_info(w), ref.vb_sqrt_info(w)),
              rel_err(out.rp_sqrt_info(w), ref.rp_sqrt_info(w)), rel_err(out.se3_sqrt_info(w), ref.se3_sqrt_info(w)), rel_err(out.pg_sqrt_info(w), ref.pg_sqrt_info(w))))
