"""GPU bring-up probe: engine variants vs the CPU oracle on a tiny network (run under gpurun)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dream_go_b200 import nn, weights   # noqa: E402
from oracle import oracle               # noqa: E402


def report(tag, got, want):
    got = np.asarray(got, np.float32)
    want = np.asarray(want, np.float32)
    d = np.abs(got - want)
    print(f"  {tag:28s} max|d|={d.max():.5f} mean|d|={d.mean():.6f} max|want|={np.abs(want).max():.3f} "
          f"nan={int(np.isnan(got).sum())}", flush=True)


def main():
    variants = sys.argv[1:] or ["direct", "tc", "nopdl"]
    nb, batch = 2, 3
    net = weights.synthetic_network(seed=7, num_blocks=nb, gate="random")
    feats = weights.bernoulli_features(batch, seed=3)
    onet = oracle.OracleNetwork(net)
    t0 = time.time()
    ov, op, oblocks = onet.forward(feats, want_blocks=True)
    print(f"oracle: {time.time() - t0:.2f}s", flush=True)
    flags = {"direct": nn.FLAG_DEBUG_DIRECT_CONV, "tc": 0, "nopdl": nn.FLAG_NO_PDL, "layerwise": nn.FLAG_LAYERWISE, "norot": nn.FLAG_NO_ROTATE}
    for v in variants:
        print(f"== {v}", flush=True)
        try:
            eng = nn.Network.from_tensors(net, max_batch=8, num_workspaces=1, flags=flags[v])
            with eng.get_workspace(batch) as ws:
                value, policy = nn.forward(ws, feats).unwrap()
            report("value", value, ov)
            report("policy", policy.reshape(batch, 362), op)
            for layer in range(nb + 1):
                report(f"tower after layer {layer}", eng.debug_read_tower(layer, batch), oblocks[layer])
            eng.close()
        except Exception as exc:   # noqa: BLE001
            print(f"  FAILED: {exc!r}", flush=True)


if __name__ == "__main__":
    main()
