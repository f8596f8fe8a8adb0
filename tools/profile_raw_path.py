#!/usr/bin/env python
"""Exercises the raw-position path for the profiler: 256 fixture-game positions through dg_engine_forward_raw_prior with the
ladders read on the device (planes_from_stones_kernel incl. the ladder reader, tower, policy FC, finish, prior kernel) and
through dg_engine_forward_packed (pack_compact_kernel).    ncu ... python tools/profile_raw_path.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dream_go_b200 import go as pgo, nn, weights
from oracle import go as ogo

raws, packed = [], []
for colors, moves, komi in ogo.load_games()[2:8]:
    b = pgo.Board(komi)
    for ply, (c, m) in enumerate(zip(colors, moves)):
        if 60 <= ply < 110 and len(raws) < 256:
            raws.append(b.raw_position(int(c), (ply % 8) | 0x08, search=ply % 2)[0])     # device ladders, both search options
            packed.append(b.features_packed(int(c), ply % 8)[0])
        if m < 361:
            b.place_index(int(c), int(m))
raws, packed = np.array(raws, nn.RAW_DTYPE), np.array(packed, nn.PACKED_DTYPE)
assert len(raws) == 256
net = nn.Network.from_tensors(weights.synthetic_network(seed=20261017, num_blocks=9), max_batch=256, num_workspaces=1)
for _ in range(4):
    out, legal, prior = net.forward_raw_prior(raws)
    ref = net.forward_packed(packed)
assert (out.value.view(np.uint16) == ref.value.view(np.uint16)).all()
net.close()
print("ok")
