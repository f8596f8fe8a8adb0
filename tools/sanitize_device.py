#!/usr/bin/env python
"""A small pass over every device code path for compute-sanitizer (memcheck / racecheck): 2-block network, 24 positions
of a fixture game -- fp16 features, compact positions, raw positions with priors and with the ladders read on the device,
the leaf-batch queue (eager first submit, captured graph, replay).
    compute-sanitizer --tool memcheck python tools/sanitize_device.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dream_go_b200 import go as pgo, nn, weights
from oracle import go as ogo

colors, moves, komi = ogo.load_games()[4]
b = pgo.Board(komi)
raws, raws_dev, packed = [], [], []
for ply, (c, m) in enumerate(zip(colors, moves)):
    if 70 <= ply < 94:
        raws.append(b.raw_position(int(c), ply % 8, search=ply % 2)[0])
        raws_dev.append(b.raw_position(int(c), (ply % 8) | 0x08, search=ply % 2)[0])
        packed.append(b.features_packed(int(c), ply % 8)[0])
    if m < 361:
        b.place_index(int(c), int(m))
raws, raws_dev, packed = np.array(raws, nn.RAW_DTYPE), np.array(raws_dev, nn.RAW_DTYPE), np.array(packed, nn.PACKED_DTYPE)
net = nn.Network.from_tensors(weights.synthetic_network(seed=7, num_blocks=2, gate="random"), max_batch=32, num_workspaces=2)
feats = pgo.unpack_features(packed)
with net.get_workspace(24) as ws:
    v0, p0 = nn.forward(ws, np.ascontiguousarray(feats)).unwrap()
o1 = net.forward_packed(packed)
o2, legal2, prior2 = net.forward_raw_prior(raws)
o3, legal3, prior3 = net.forward_raw_prior(raws_dev)
assert (o1.value.view(np.uint16) == v0.view(np.uint16)).all() and (o2.value.view(np.uint16) == v0.view(np.uint16)).all()
assert (o3.policy.view(np.uint16) == o2.policy.view(np.uint16)).all() and (prior3.view(np.uint32) == prior2.view(np.uint32)).all()
with net.leaf_batch() as lb:
    for rnd in range(3):
        assert lb.push(raws_dev) == 0
        lb.submit(prior=True)
        lb.wait()
        v, p, lg, pr = lb.results(prior=True)
        assert (pr.view(np.uint32) == prior2.view(np.uint32)).all() and (lg == legal2).all()
        lb.reset()
net.close()
print("sanitize_device: ok")
