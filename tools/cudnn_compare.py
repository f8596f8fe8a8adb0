"""Engine vs the cuDNN restatement of the reference forward vs the CPU oracle, plus timing (GPU box)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dream_go_b200 import nn, weights
from baseline import cudnn_ref
from oracle import oracle

def stats(tag, got, want):
    d = np.abs(np.asarray(got, np.float32) - np.asarray(want, np.float32))
    print(f"  {tag:34s} max|d|={d.max():.5f} mean|d|={d.mean():.7f}", flush=True)

net = weights.synthetic_network(seed=20261017, num_blocks=9)
b = 16
feats = weights.bernoulli_features(b, seed=3)
ov, op, ot = oracle.OracleNetwork(net).forward(feats, want_tower=True)
ref = cudnn_ref.CudnnNetwork(net, b)
print("cuDNN:", ref.info)
cv, cp = ref.forward(feats)
ct = ref.read_tower()
eng = nn.Network.from_tensors(net, max_batch=256, num_workspaces=1)
with eng.get_workspace(b) as ws:
    ev, ep = nn.forward(ws, feats).unwrap()
et = eng.debug_read_tower(-1, b)
print("vs CPU oracle (fp64 accumulate, fp16 storage):")
stats("cuDNN value", cv, ov); stats("engine value", ev, ov)
stats("cuDNN policy", cp, op); stats("engine policy", ep.reshape(b, 362), op)
stats("cuDNN tower", ct, ot); stats("engine tower", et, ot)
print("top-1 agreement vs oracle: cuDNN %.3f engine %.3f" % ((cp.argmax(1) == op.argmax(1)).mean(), (ep.reshape(b, 362).argmax(1) == op.argmax(1)).mean()))
ref.close()
for batch in (1, 8, 16, 32, 64, 128, 256):
    f = weights.bernoulli_features(batch, seed=batch)
    r = cudnn_ref.CudnnNetwork(net, batch)
    r.forward(f); r.time_resident(5)
    iters = 50
    ms_ref = r.time_resident(iters) / iters
    t0 = time.perf_counter()
    for _ in range(iters): r.forward(f)
    e2e_ref = (time.perf_counter() - t0) / iters * 1e3
    r.close()
    v = np.empty(batch, np.float16); p = np.empty((batch, 362), np.float16)
    eng.forward_into(f, v, p)
    ms_eng = eng.time_resident(batch, iters, tower=False, flush_l2=False)[0] / iters
    t0 = time.perf_counter()
    for _ in range(iters): eng.forward_into(f, v, p)
    e2e_eng = (time.perf_counter() - t0) / iters * 1e3
    print(f"batch {batch:4d}: cuDNN {ms_ref:.3f} ms resident / {e2e_ref:.3f} ms e2e   engine {ms_eng:.3f} ms resident / {e2e_eng:.3f} ms e2e   speedup {ms_ref/ms_eng:.2f}x / {e2e_ref/e2e_eng:.2f}x", flush=True)
