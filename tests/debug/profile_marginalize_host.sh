# ncu launch list of one C++ MarginalizationInfo chain (two marginalize() rounds + 6 warm repeats): which kernels make up
# the host latency of isv_marginalize_host.  Run on the GPU box: bash tests/debug/profile_marginalize_host.sh
set -e
cd "$(dirname "$0")/../.."
g++ -std=c++17 -O2 -o build/marginalization_chain_test tests/cpp/marginalization_chain_test.cpp -Lis_vins_b200 -lisv_b200 -Wl,-rpath,$PWD/is_vins_b200
python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
from tests.test_host_cpp_gpu import _chain_problem
from oracle import sim
p = sim.make_problem(sim.seed_for(9, 21), n_features=320, max_track=9, host0=0.4)
d, rounds = _chain_problem(p, False, 77)
np.asarray(d, dtype="<f8").tofile("gpurun_out/chain_fx.bin")
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s20_chain_launches.csv build/marginalization_chain_test gpurun_out/chain_fx.bin gpurun_out/chain_dump.bin > gpurun_out/s20_chain.log 2>&1
tail -2 gpurun_out/s20_chain.log
