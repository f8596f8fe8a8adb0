#!/usr/bin/env python
"""Same-box A/B of engine flags (one process): resident forward time at batch 256, ~1 s of device time per measurement,
alternating.    python tools/ab_flags.py 0 128      (flag words, see include/dg_engine.h)"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dream_go_b200 import nn, weights
t = weights.synthetic_network(seed=20261017, num_blocks=9)
out = []
for rnd in range(2):
    for flags in [int(a, 0) for a in sys.argv[1:]]:
        net = nn.Network.from_tensors(t, max_batch=256, num_workspaces=1, flags=flags)
        f = net.pinned((256, 361, 32), np.float16); f[...] = weights.bernoulli_features(256, seed=3)
        v, p = net.pinned((256,), np.float16), net.pinned((256, 362), np.float16)
        for _ in range(5): net.forward_into(f, v, p)
        ms, tms, launches = net.time_resident(256, 2500, tower=True, flush_l2=True)
        out.append({"flags": flags, "round": rnd, "ms_forward": ms / 2500, "us_tower_18_convs": 1e3 * tms / 2500, "launches": launches,
                    "checksum": float(p.astype(np.float64).sum()), "policy_crc": int(np.bitwise_xor.reduce(p.view(np.uint16).astype(np.uint64).ravel() * np.arange(1, p.size + 1, dtype=np.uint64)))})
        net.close()
print(json.dumps(out))
