#!/usr/bin/env python
"""Same-box A/B of two builds of the engine library: resident forward time and tower-only time at batch 256 (sustained:
~1 s of device time per measurement), alternating A, B, A, B.

    python tools/ab_resident.py build/libdg_engine_r01.so dream_go_b200/libdg_engine.so
Each library is loaded in its own process (DG_ENGINE_LIB)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import numpy as np
from dream_go_b200 import nn, weights
t = weights.synthetic_network(seed=20261017, num_blocks=9)
B = int(os.environ.get("DG_AB_BATCH", "256"))
net = nn.Network.from_tensors(t, max_batch=B, num_workspaces=1)
f = net.pinned((B, 361, 32), np.float16); f[...] = weights.bernoulli_features(B, seed=3)
v, p = net.pinned((B,), np.float16), net.pinned((B, 362), np.float16)
for _ in range(5): net.forward_into(f, v, p)
iters = 2500 if B >= 256 else 4000
ms, tms, _ = net.time_resident(B, iters, tower=True, flush_l2=True)
print(json.dumps({"batch": B, "ms_forward": ms / iters, "us_tower": 1e3 * tms / iters, "checksum": float(p.astype(np.float64).sum())}))
''' % ROOT
out = []
for rnd in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, DG_ENGINE_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"error": r.stderr[-300:]})
        out.append({"lib": lib, "round": rnd, **json.loads(line)})
print(json.dumps(out))
