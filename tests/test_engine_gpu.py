"""Parity of the sm_100a engine (through the C ABI) against the CPU oracle of dg_nn::forward.

Tolerances: value |d| <= 4e-3; policy |d| <= 1e-3 + 1e-2 * p; tower
activations: relative L2 error <= 2e-3 (fp32 tensor-core accumulation vs the oracle's double
accumulation, both rounded to fp16 after every layer)."""
import threading

import numpy as np
import pytest

from dream_go_b200 import nn, weights
from oracle import oracle

pytestmark = pytest.mark.gpu

VALUE_TOL = 4e-3


def check_outputs(value, policy, want_v, want_p):
    v, wv = value.astype(np.float32), want_v.astype(np.float32)
    p, wp = policy.reshape(-1, 362).astype(np.float32), want_p.astype(np.float32)
    assert np.isfinite(v).all() and np.isfinite(p).all()
    assert np.abs(v - wv).max() <= VALUE_TOL
    assert (np.abs(p - wp) <= 1e-3 + 1e-2 * wp).all()
    assert np.abs(p.sum(axis=1) - 1.0).max() < 1e-2


def rel_l2(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30))


@pytest.fixture(scope="module")
def small_engine(small_net):
    net = nn.Network.from_tensors(small_net, max_batch=64, num_workspaces=2)
    yield net
    net.close()


@pytest.fixture(scope="module")
def full_engine(full_net):
    net = nn.Network.from_tensors(full_net, max_batch=256, num_workspaces=2)
    yield net
    net.close()


@pytest.mark.parametrize("batch", [1, 2, 7, 33, 64])
def test_small_network_matches_oracle(small_engine, small_net, batch):
    feats = weights.bernoulli_features(batch, seed=100 + batch)
    with small_engine.get_workspace(batch) as ws:
        value, policy = nn.forward(ws, feats).unwrap()
    want_v, want_p, blocks = oracle.OracleNetwork(small_net).forward(feats, want_blocks=True)
    check_outputs(value, policy, want_v, want_p)
    for layer in range(small_engine.num_blocks + 1):
        got = small_engine.debug_read_tower(layer, batch)
        assert rel_l2(got, blocks[layer]) <= 2e-3, layer
        assert np.abs(got.astype(np.float32) - blocks[layer].astype(np.float32)).max() <= 2e-2


def test_full_network_matches_oracle(full_engine, full_net):
    batch = 12
    feats = weights.bernoulli_features(batch, seed=9)
    with full_engine.get_workspace(batch) as ws:
        value, policy = nn.forward(ws, feats).unwrap()
    want_v, want_p, tower = oracle.OracleNetwork(full_net).forward(feats, want_tower=True)
    check_outputs(value, policy, want_v, want_p)
    assert rel_l2(full_engine.debug_read_tower(-1, batch), tower) <= 2e-3
    assert full_engine.num_blocks == 9


@pytest.mark.parametrize("inputs", ["bernoulli", "real"])
def test_full_network_batch_256_matches_oracle(full_engine, full_net, inputs):
    """BASELINE.json configs[1] at its stated size -- 9 blocks x 128 filters, batch 256 -- against the CPU oracle directly
    (value, policy, final tower activation), on the reference benchmark's Bernoulli(0.2) inputs
    (dg_tests/benches/batch_sizes.rs:44-49) and on real V1 feature planes of fixture-game positions."""
    batch = 256
    if inputs == "bernoulli":
        feats = weights.bernoulli_features(batch, seed=256)
    else:
        from oracle import go as ogo
        games = ogo.load_games()
        long_games = [g for g in games if len(g[1]) >= 160][:4]
        feats = np.concatenate([ogo.replay(c[:160], m[:160], k, features=True)["features"][96:160] for c, m, k in long_games])[:batch]
        assert feats.shape == (batch, 361, 32)
    with full_engine.get_workspace(batch) as ws:
        value, policy = nn.forward(ws, np.ascontiguousarray(feats)).unwrap()
    want_v, want_p, tower = oracle.OracleNetwork(full_net).forward(feats, want_tower=True)
    check_outputs(value, policy, want_v, want_p)
    got = full_engine.debug_read_tower(-1, batch)
    assert rel_l2(got, tower) <= 2e-3
    per_pos = [rel_l2(got[i], tower[i]) for i in range(batch)]       # no position of the batch is off (tile / pair boundaries)
    assert max(per_pos) <= 4e-3, int(np.argmax(per_pos))


def test_random_gates_full_depth():
    tensors = weights.synthetic_network(seed=5, num_blocks=9, gate="random")
    feats = weights.bernoulli_features(4, seed=10)
    net = nn.Network.from_tensors(tensors, max_batch=4)
    with net.get_workspace(4) as ws:
        value, policy = nn.forward(ws, feats).unwrap()
    want_v, want_p = oracle.OracleNetwork(tensors).forward(feats)
    check_outputs(value, policy, want_v, want_p)
    net.close()


def test_edge_inputs(small_engine, small_net):
    """All-zero planes, all-one planes and a lone stone in each corner (padding / halo handling)."""
    feats = np.zeros((6, 361, 32), np.float16)
    feats[1] = 1.0
    for i, point in enumerate([0, 18, 342, 360]):
        feats[2 + i, point, :] = 1.0
    with small_engine.get_workspace(6) as ws:
        value, policy = nn.forward(ws, feats).unwrap()
    want_v, want_p, blocks = oracle.OracleNetwork(small_net).forward(feats, want_blocks=True)
    check_outputs(value, policy, want_v, want_p)
    got = small_engine.debug_read_tower(0, 6)
    assert np.abs(got.astype(np.float32) - blocks[0].astype(np.float32)).max() <= 4e-3


def test_tensor_core_kernel_matches_direct_kernel_at_full_size(full_net):
    """BASELINE size (batch 256, 9 blocks): tcgen05 path vs the one-thread-per-output cross-check
    kernel on the device (the CPU oracle takes minutes at this size)."""
    feats = weights.bernoulli_features(256, seed=2)
    outs = []
    for flags in (0, nn.FLAG_DEBUG_DIRECT_CONV):
        net = nn.Network.from_tensors(full_net, max_batch=256, num_workspaces=1, flags=flags)
        with net.get_workspace(256) as ws:
            value, policy = nn.forward(ws, feats).unwrap()
        outs.append((value.copy(), policy.copy(), net.debug_read_tower(-1, 256)))
        net.close()
    (v0, p0, t0), (v1, p1, t1) = outs
    assert rel_l2(t0, t1) <= 1e-3
    assert np.abs(v0.astype(np.float32) - v1.astype(np.float32)).max() <= VALUE_TOL
    assert np.abs(p0.astype(np.float32) - p1.astype(np.float32)).max() <= 1e-3
    # size-independent property: every position of the batch is evaluated independently
    net = nn.Network.from_tensors(full_net, max_batch=256, num_workspaces=1)
    sel = [0, 100, 255]
    with net.get_workspace(3) as ws:
        v3, p3 = nn.forward(ws, feats[sel]).unwrap()
    assert np.array_equal(v3, v0[sel])
    assert np.array_equal(p3.reshape(3, 362), p0.reshape(256, 362)[sel])
    net.close()


def test_batch_shrink_leaves_no_stale_rows(small_engine, small_net):
    big = weights.bernoulli_features(64, seed=1)
    small = weights.bernoulli_features(3, seed=2)
    with small_engine.get_workspace(64) as ws:
        nn.forward(ws, big)
    with small_engine.get_workspace(64) as ws:
        nn.forward(ws, big)            # touch both pooled workspaces
    with small_engine.get_workspace(3) as ws:
        value, policy = nn.forward(ws, small).unwrap()
    want_v, want_p = oracle.OracleNetwork(small_net).forward(small)
    check_outputs(value, policy, want_v, want_p)


def test_deterministic_and_packed_path_bit_exact(small_engine):
    feats = weights.bernoulli_features(17, seed=4)
    feats[:, :, 0] = np.float16(0.9667)         # k plane (to move = black)
    feats[:, :, 1] = 0
    with small_engine.get_workspace(17) as ws:
        v1, p1 = nn.forward(ws, feats).unwrap()
        v2, p2 = nn.forward(ws, feats).unwrap()
    assert np.array_equal(v1, v2) and np.array_equal(p1, p2)
    out = small_engine.forward_packed(nn.pack_positions(feats))
    assert np.array_equal(out.value, v1)
    assert np.array_equal(out.policy.reshape(-1), p1)


def _raw_positions(n, game=4, start=20):
    """n raw positions (consecutive plies of a fixture game, cycling through the symmetries and both search options)."""
    from dream_go_b200 import go as pgo
    from oracle import go as ogo
    colors, moves, komi = ogo.load_games()[game]
    board = pgo.Board(komi)
    out = []
    for ply, (c, m) in enumerate(zip(colors, moves)):
        if ply >= start and len(out) < n:
            out.append(board.raw_position(int(c), (ply % 8) | ((ply // 8 % 2) << 4)))
        if m < 361:
            board.place_index(int(c), int(m))
    assert len(out) == n
    return np.concatenate(out)


@pytest.mark.parametrize("flags", [0, nn.FLAG_NO_GRAPH])
def test_leaf_batch_queue_matches_blocking_forward(small_net, flags):
    """The leaf-batch queue (lock-free pushes from several threads, one captured graph launch per submit, completion flag in
    pinned memory) returns bit for bit what dg_engine_forward_raw_prior returns -- first submit (eager), replays of a captured
    bucket, a second bucket, a batch that fills the workspace, and with the capture switched off."""
    net = nn.Network.from_tensors(small_net, max_batch=64, num_workspaces=2, flags=flags)
    try:
        raw = _raw_positions(150)
        want, want_legal, want_prior = net.forward_raw_prior(raw[:64])
        _, _, pa = net.forward_raw_prior(raw[64:128])
        _, _, pb = net.forward_raw_prior(raw[128:150])
        with net.leaf_batch() as lb, net.leaf_batch() as lb2:        # both workspaces of the engine: a blocking forward would wait now
            assert lb.capacity == 64
            for rnd, (lo, hi, prior) in enumerate([(0, 37, True), (0, 37, True), (3, 40, True), (0, 64, True), (10, 15, False), (0, 64, False)]):
                order = [None] * (hi - lo)

                def producer(k, lo=lo, hi=hi, order=order):
                    for i in range(lo + k, hi, 3):
                        order[i - lo] = lb.push(raw[i:i + 1])

                threads = [threading.Thread(target=producer, args=(k,)) for k in range(3)]
                [t.start() for t in threads]
                [t.join() for t in threads]
                assert sorted(order) == list(range(hi - lo))
                if hi - lo == 64:
                    assert lb.push(raw[:1]) == -1              # full
                lb.submit(prior=prior)
                assert lb.push(raw[:1]) == -1                  # sealed until reset
                lb.wait()
                assert lb.ready()
                res = lb.results(prior=prior)
                slot = np.array(order)
                idx = np.arange(lo, hi)
                assert (res[0][slot].view(np.uint16) == want.value[idx].view(np.uint16)).all(), rnd
                assert (res[1][slot].view(np.uint16) == want.policy.reshape(-1, 362)[idx].view(np.uint16)).all(), rnd
                assert (res[2][slot] == want_legal[idx]).all(), rnd
                if prior:
                    assert (res[3][slot].view(np.uint32) == want_prior[idx].view(np.uint32)).all(), rnd
                lb.reset()
            # two batches in flight at once, as the reference allows per device (predictors/nn.rs:64-67)
            assert lb.push(raw[64:100]) == 0 and lb2.push(raw[100:150]) == 0
            lb.submit(prior=True)
            lb2.submit(prior=True)
            lb2.wait()
            lb.wait()
            got = np.concatenate([lb.results(True)[3], lb2.results(True)[3]])
            assert (got.view(np.uint32) == np.concatenate([pa, pb]).view(np.uint32)).all()
            with pytest.raises(nn.Error):
                lb.submit()                                    # already submitted, not reset
    finally:
        net.close()


def test_concurrent_forwards(small_engine, small_net):
    """The reference allows 2 forwards in flight per device (predictors/nn.rs:64-67)."""
    onet = oracle.OracleNetwork(small_net)
    errors = []

    def worker(seed):
        try:
            feats = weights.bernoulli_features(5 + seed, seed=seed)
            for _ in range(5):
                with small_engine.get_workspace(5 + seed) as ws:
                    value, policy = nn.forward(ws, feats).unwrap()
            want_v, want_p = onet.forward(feats)
            check_outputs(value, policy, want_v, want_p)
        except Exception as exc:   # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors


def test_device_helpers():
    import torch
    L = nn.lib()
    assert L.dg_device_count() == torch.cuda.device_count() >= 1      # every device of a B200 box is an sm_100 part
    assert L.dg_set_current_device(0) == 0 and L.dg_current_device() == 0
    assert L.dg_set_current_device(L.dg_device_count()) == -1


def test_one_process_two_devices(small_net):
    """The reference drives every GPU from ONE process (`cudaSetDevice` per call, predictors/nn.rs:84-92): two engines on
    two devices, used concurrently from two threads, give the results of one engine alone, bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nets = [nn.Network.from_tensors(small_net, device=d, max_batch=64, num_workspaces=2) for d in (1, 0)]
    feats = [weights.bernoulli_features(40 + d, seed=70 + d) for d in range(2)]
    alone = []
    for net, f in zip(nets, feats):
        with net.get_workspace(len(f)) as ws:
            alone.append(nn.forward(ws, f).unwrap())
    got, errors = [None, None], []

    def worker(i):
        try:
            for _ in range(20):
                with nets[i].get_workspace(len(feats[i])) as ws:
                    got[i] = nn.forward(ws, feats[i]).unwrap()
        except Exception as exc:   # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for i in range(2):
        assert np.array_equal(got[i][0], alone[i][0]) and np.array_equal(got[i][1], alone[i][1])
    want_v, want_p = oracle.OracleNetwork(small_net).forward(feats[0])
    check_outputs(*got[0], want_v, want_p)
    for net in nets:
        net.close()


def test_errors(tmp_path, small_net):
    net = nn.Network(max_batch=4)
    with pytest.raises(nn.Error) as err:      # forward before load
        net.forward_into(np.zeros((1, 361, 32), np.float16), np.zeros(1, np.float16), np.zeros((1, 362), np.float16))
    assert err.value.kind == "MissingWeights"
    with pytest.raises(nn.Error) as err:
        net.load_json(str(tmp_path / "missing.json"))
    assert err.value.kind == "MissingWeights"
    bad = tmp_path / "bad.json"
    bad.write_text('{"01_upsample/conv_1:0": {"t": "zz", "v": "00000"}}')
    with pytest.raises(nn.Error) as err:
        net.load_json(str(bad))
    assert err.value.kind == "MalformedWeights"
    incomplete = dict(small_net)
    del incomplete["01_upsample/conv_1/offset:0"]
    with pytest.raises(nn.Error) as err:
        net.load_tensors(incomplete)
    assert err.value.kind == "MissingWeights"
    with pytest.raises(nn.Error):
        net.get_workspace(5)
    path = tmp_path / "dream_go.json"
    weights.dump_json(small_net, str(path))
    net.load_json(str(path))                  # the JSON route gives the same network as raw tensors
    feats = weights.bernoulli_features(2, seed=8)
    with net.get_workspace(2) as ws:
        value, policy = nn.forward(ws, feats).unwrap()
    want_v, want_p = oracle.OracleNetwork(oracle.load_json(str(path))).forward(feats)
    check_outputs(value, policy, want_v, want_p)
    net.close()


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


def test_no_further_from_oracle_than_cudnn(full_engine, full_net):
    """"the new engine must be no further from oracle 2 than cuDNN is" -- the cuDNN
    restatement of the reference forward (baseline/cudnn_ref.cu) is the second oracle."""
    from baseline import cudnn_ref
    batch = 8
    feats = weights.bernoulli_features(batch, seed=77)
    want_v, want_p, want_t = oracle.OracleNetwork(full_net).forward(feats, want_tower=True)
    ref = cudnn_ref.CudnnNetwork(full_net, batch)
    cv, cp = ref.forward(feats)
    ct = ref.read_tower()
    ref.close()
    with full_engine.get_workspace(batch) as ws:
        ev, ep = nn.forward(ws, feats).unwrap()
    et = full_engine.debug_read_tower(-1, batch)
    check_outputs(cv, cp, want_v, want_p)                      # the baseline itself agrees with the CPU oracle
    assert rel_l2(et, want_t) <= max(rel_l2(ct, want_t), 5e-4)
    assert np.abs(ev.astype(np.float32) - cv.astype(np.float32)).max() <= VALUE_TOL
    assert np.abs(ep.reshape(batch, 362).astype(np.float32) - cp.astype(np.float32)).max() <= 1e-3


@pytest.mark.parametrize("batch", [300, 512, 1000])
def test_large_and_ragged_batches_equal_split_evaluation(full_net, batch):
    """BASELINE-size network at batches beyond 256 (self-play sends up to 512 leaves per forward): every position's
    output is bit-identical to its output in a batch of 64 (batch-position independence at full size)."""
    net = nn.Network.from_tensors(full_net, max_batch=1024, num_workspaces=2)
    try:
        feats = weights.bernoulli_features(batch, seed=batch)
        with net.get_workspace(batch) as ws:
            value, policy = nn.forward(ws, feats).unwrap()
        policy = policy.reshape(batch, 362)
        assert np.isfinite(policy.astype(np.float32)).all()
        assert np.abs(policy.astype(np.float32).sum(axis=1) - 1.0).max() < 1e-2
        for at in (0, batch // 2 - 17, batch - 64):
            with net.get_workspace(64) as ws:
                v, p = nn.forward(ws, feats[at:at + 64]).unwrap()
            assert (v.view(np.uint16) == value[at:at + 64].view(np.uint16)).all()
            assert (p.reshape(64, 362).view(np.uint16) == policy[at:at + 64].view(np.uint16)).all()
    finally:
        net.close()
