#!/bin/bash
# tests of the device ladder reader + position-aligned tower tiles, then A/B of the producer order (trace + timing)
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02d_pytest.log; tail -4 gpurun_out/r02d_pytest.log
python tools/tower_trace.py 256 0 > gpurun_out/r02d_trace_early_a.txt 2>&1; cp gpurun_out/tower_trace.npy gpurun_out/r02d_trace_early_a.npy
python tools/tower_trace.py 256 64 > gpurun_out/r02d_trace_late_a.txt 2>&1
head -4 gpurun_out/r02d_trace_early_a.txt; head -4 gpurun_out/r02d_trace_late_a.txt
python - <<'PY' > gpurun_out/r02d_ab.json
import json, numpy as np
from dream_go_b200 import nn, weights
t = weights.synthetic_network(seed=20261017, num_blocks=9)
out = {}
for name, flags in (("early_a", 0), ("late_a", 64), ("early_a_again", 0), ("late_a_again", 64)):
    net = nn.Network.from_tensors(t, max_batch=256, num_workspaces=1, flags=flags)
    f = net.pinned((256, 361, 32), np.float16); f[...] = weights.bernoulli_features(256, seed=3)
    v, p = net.pinned((256,), np.float16), net.pinned((256, 362), np.float16)
    for _ in range(5): net.forward_into(f, v, p)
    ms, tms, _ = net.time_resident(256, 2500, tower=True, flush_l2=True)
    out[name] = {"ms_forward": ms / 2500, "us_tower": 1e3 * tms / 2500}
    net.close()
print(json.dumps(out))
PY
cat gpurun_out/r02d_ab.json
