"""Prints the in-kernel clock64() timeline of one tower convolution launch (perf debugging)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dream_go_b200 import nn, weights

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = nn.Network.from_tensors(weights.synthetic_network(num_blocks=2), max_batch=batch, num_workspaces=1)
feats = weights.bernoulli_features(batch, seed=1)
with net.get_workspace(batch) as ws:
    nn.forward(ws, feats)
tr = net.debug_conv_trace(batch)
tr = net.debug_conv_trace(batch)
for cta in (0, 1, 73, 147):
    t = tr[cta]
    t0 = t[t > 0].min()
    print(f"cta {cta}")
    for role, name in enumerate(["producer", "mma", "epilogue"]):
        v = t[role][t[role] > 0] - t0
        print(f"  {name:9s}", " ".join(str(int(x)) for x in v))
start = np.where(tr > 0, tr, np.iinfo(np.int64).max).min(axis=(1, 2))
end = tr.max(axis=(1, 2))
print("per-CTA duration cycles: min %d median %d max %d" % ((end - start).min(), np.median(end - start), (end - start).max()))
