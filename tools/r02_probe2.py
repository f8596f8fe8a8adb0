#!/usr/bin/env python
"""Round-2 probe: resident forward time by batch size beyond 256 (the layer-boundary cost is fixed per layer, the work is not)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dream_go_b200 import nn, weights

tensors = weights.synthetic_network(seed=20261017, num_blocks=9)
rows = []
for batch in (128, 256, 384, 512, 768, 1024):
    net = nn.Network.from_tensors(tensors, max_batch=batch, num_workspaces=1)
    feats = net.pinned((batch, 361, 32), np.float16)
    feats[...] = weights.bernoulli_features(batch, seed=batch)
    value, policy = net.pinned((batch,), np.float16), net.pinned((batch, 362), np.float16)
    for _ in range(5):
        net.forward_into(feats, value, policy)
    iters = max(200, int(1.0 / (batch * 1.6e-6)))
    ms, tms, _ = net.time_resident(batch, iters, tower=True, flush_l2=True)
    net.close()
    rows.append({"batch": batch, "iters": iters, "ms_forward": ms / iters, "ms_tower_18_convs": tms / iters, "evals_per_s": batch * iters / (ms * 1e-3),
                 "us_per_position": 1e3 * ms / iters / batch, "tower_tflops": 18 * batch * 2 * 361 * 1152 * 128 / (tms / iters * 1e-3) / 1e12})
print(json.dumps({"metric": "resident_forward_by_batch", "rows": rows}))
