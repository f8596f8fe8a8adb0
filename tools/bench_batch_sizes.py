#!/usr/bin/env python
"""The reference's own micro-benchmark `dg_tests/benches/batch_sizes.rs:41-68` (forward of B positions of iid
Bernoulli(0.2) features, B in {1, 8, 16, 32, 64, 128, 256}) on this engine and on the cuDNN restatement of the reference,
same box, same weights: device-resident ms per forward (CUDA events, L2 flushed) and the blocking host-to-host call.

    python tools/bench_batch_sizes.py        -> one JSON line
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from baseline import cudnn_ref
    from dream_go_b200 import nn, weights
    tensors = weights.synthetic_network(seed=20261017, num_blocks=9)
    rows = []
    for batch in (1, 8, 16, 32, 64, 128, 256):
        feats_src = weights.bernoulli_features(batch, seed=batch)
        net = nn.Network.from_tensors(tensors, max_batch=batch, num_workspaces=1)
        feats = net.pinned((batch, 361, 32), np.float16)
        feats[...] = feats_src
        value, policy = net.pinned((batch,), np.float16), net.pinned((batch, 362), np.float16)
        for _ in range(5):
            net.forward_into(feats, value, policy)
        ms, _, _ = net.time_resident(batch, 200, tower=False, flush_l2=True)
        t0 = time.perf_counter()
        for _ in range(200):
            net.forward_into(feats, value, policy)
        call = (time.perf_counter() - t0) / 200
        net.close()
        ref = cudnn_ref.CudnnNetwork(tensors, batch)
        host = np.array(feats_src)
        ref.forward(host)
        ref.time_resident(5)
        ref_ms = ref.time_resident(100) / 100
        t0 = time.perf_counter()
        for _ in range(50):
            ref.forward(host)
        ref_call = (time.perf_counter() - t0) / 50
        ref.close()
        rows.append({"batch": batch, "engine_ms_resident": ms / 200, "engine_ms_call": call * 1e3, "cudnn_ms_resident": ref_ms,
                     "cudnn_ms_call": ref_call * 1e3, "speedup_resident": ref_ms / (ms / 200), "speedup_call": ref_call / call})
    print(json.dumps({"metric": "forward_ms_by_batch_size", "unit": "ms", "higher_is_better": False,
                      "config": {"workload": "dg_tests/benches/batch_sizes.rs: 9 blocks x 128 filters, iid Bernoulli(0.2) features"},
                      "rows": rows}))


if __name__ == "__main__":
    main()
