"""Timeline statistics of one persistent tower launch from in-kernel clock64() samples (perf debugging)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dream_go_b200 import nn, weights

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
net = nn.Network.from_tensors(weights.synthetic_network(num_blocks=9), max_batch=batch, num_workspaces=1, flags=flags)
feats = weights.bernoulli_features(batch, seed=1)
with net.get_workspace(batch) as ws:
    nn.forward(ws, feats)
net.debug_tower_trace(batch)
tr = net.debug_tower_trace(batch)
np.save("gpurun_out/tower_trace.npy", tr)
tot = []
for cta in range(0, 148, 2):
    mma = tr[cta, 1]; mma = mma[mma > 0]
    n = len(mma) // 4
    ev = mma[:n * 4].reshape(n, 4)          # unit start, acc free, operands landed, issued
    tot.append((ev[-1, 3] - ev[0, 0], n, (ev[:, 1] - ev[:, 0]).sum(), (ev[:, 2] - ev[:, 1]).sum(), (ev[:, 3] - ev[:, 2]).sum()))
tot = np.array(tot, dtype=np.float64)
print("pairs: units/pair min %d max %d" % (tot[:, 1].min(), tot[:, 1].max()))
print("MMA issuer per pair (cycles): span %.0f | wait accumulator %.0f | wait operands(k-half 0)+weights %.0f | issue+k-half-1 wait %.0f"
      % (tot[:, 0].mean(), tot[:, 2].mean(), tot[:, 3].mean(), tot[:, 4].mean()))
print("  per unit: span %.0f  acc-wait %.0f  operand-wait %.0f  issue %.0f" % tuple((tot[:, [0, 2, 3, 4]].sum(0) / tot[:, 1].sum())))
cta = 0
mma = tr[cta, 1]; mma = mma[mma > 0]; n = len(mma) // 4; ev = mma[:n * 4].reshape(n, 4); ev = ev - ev[0, 0]
print("cta0 first 14 units (start, acc free, operands, issued):")
for i in range(14): print("   ", ev[i].tolist())
pr = tr[cta, 0]; pr = pr[pr > 0]
ep = tr[cta, 2]; ep = ep[ep > 0]; m = len(ep) // 3; ee = ep[:m * 3].reshape(m, 3)
print("epilogue cta0: mean wait %.0f  mean work %.0f  mean gap-to-next %.0f" % ((ee[:, 1] - ee[:, 0]).mean(), (ee[:, 2] - ee[:, 1]).mean(), (ee[1:, 0] - ee[:-1, 2]).mean()))
t0 = np.where(tr > 0, tr, np.iinfo(np.int64).max).min(); t1 = tr.max()
print("kernel span (cycles):", t1 - t0)
