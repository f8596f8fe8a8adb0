"""Pins the CPU oracle against every known-answer test the reference holds for
the NN path, and cross-checks it against an independent
fp64 torch/numpy restatement."""
import base64
import json

import numpy as np
import pytest

from oracle import oracle


# ---- src/libdg_utils/types/fp16.rs:77-92
@pytest.mark.parametrize("bits,value", [
    (0x4170, 2.71875), (0x4248, 3.140625), (0x3518, 0.31835938), (0x398c, 0.6933594),
    (0x36f3, 0.43432617), (0x3dc5, 1.4423828), (0x3da8, 1.4140625)])
def test_fp16_to_f32_kat(bits, value):
    assert oracle.f16_bits_to_f32(bits) == np.float32(value)


def test_f32_to_fp16_kat():
    assert oracle.f32_to_f16_bits(np.pi) == 0x4248
    assert oracle.f32_to_f16_bits(np.e) == 0x4170


def test_fp16_exhaustive_roundtrip_and_rne():
    # every finite half survives a round trip; conversion agrees with numpy (IEEE RNE)
    bits = np.arange(0x10000, dtype=np.uint32).astype(np.uint16)
    halves = bits.view(np.float16)
    finite = np.isfinite(halves)
    for b in bits[finite][::97]:
        assert oracle.f32_to_f16_bits(oracle.f16_bits_to_f32(int(b))) == int(b)
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.normal(0, 1, 2000), rng.normal(0, 1e-6, 500), rng.normal(0, 3e4, 500),
                         [65504.0, 65519.9, 65520.0, 1e9, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0001, 0.0, -0.0]]
                        ).astype(np.float32)
    with np.errstate(over="ignore"):
        want = xs.astype(np.float16).view(np.uint16)
    got = np.array([oracle.f32_to_f16_bits(float(x)) for x in xs], dtype=np.uint16)
    assert np.array_equal(got, want)


# ---- src/libdg_utils/b85.rs:168-220
def test_b85_pi_e():
    assert np.frombuffer(oracle.b85_decode(b"NJ4Ny"), "<f2").tolist() == [3.140625, 2.71875]


def test_b85_padding_1234567():
    got = np.frombuffer(oracle.b85_decode(b"06YLd073vn07U>s07n1-"), "<f2").tolist()
    assert got == [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 0.0]


def test_b85_f32():
    assert np.frombuffer(oracle.b85_decode(b"000<4"), "<f4").tolist() == [9.5]


def test_b85_encode_kat():
    raw = np.asarray([2.7578125, 3.67382812], dtype="<f2").tobytes()
    assert oracle.b85_encode(raw) == b"gh5$D"


def test_b85_decode_encode_roundtrip_and_python_agreement():
    examples = [[3.140625, 2.71875],
                [5.3203125, 9.9765625, 3.28320312, 8.15625, 7.109375, 1.81640625, 1.69921875, 4.4296875],
                [6.37890625, 9.6171875, 2.2890625, 9.4609375, 7.8984375, 9.3125, 4.10546875, 9.390625]]
    for ex in examples:
        raw = np.asarray(ex, dtype="<f2").tobytes()
        enc = oracle.b85_encode(raw)
        assert enc == base64.b85encode(raw)          # RFC 1924 alphabet == Python's b85
        assert oracle.b85_decode(enc) == raw
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, 4096, dtype=np.uint8).tobytes()
    assert oracle.b85_decode(base64.b85encode(raw)) == raw


def test_b85_invalid_character():
    with pytest.raises(ValueError):
        oracle.b85_decode(b"NJ4N\"")


# ---- src/libdg_nn/loader.rs:124-142
def test_loader_kat(tmp_path):
    path = tmp_path / "w.json"
    path.write_text('{"11v_value/linear_2/offset:0": {"s": "(^d>V", "t": "f2", "v": "(^d>V"}}')
    out = oracle.load_json(str(path))
    assert list(out) == ["11v_value/linear_2/offset:0"]
    assert out["11v_value/linear_2/offset:0"].nbytes == 4
    assert np.frombuffer(oracle.b85_decode(b"(^d>V"), "<f4")[0] == np.float32(0.13704996)


def test_loader_empty_is_missing(tmp_path):
    path = tmp_path / "w.json"
    path.write_text("{}")
    with pytest.raises(ValueError):
        oracle.load_json(str(path))


# ---- src/libdg_nn/layers/conv2d.rs:253-291  (pick_middle_and_sum)
def test_conv2d_pick_middle_and_sum():
    w = np.asarray([0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1], dtype=np.float16).reshape(2, 3, 3, 1)
    x = np.arange(1, 10, dtype=np.float32).reshape(1, 3, 3, 1)
    y = oracle.conv3x3(x, w, np.zeros(2, np.float16))
    assert y.reshape(-1).tolist() == [1, 12, 2, 21, 3, 16, 4, 27, 5, 45, 6, 33, 7, 24, 8, 39, 9, 28]


# ---- src/libdg_nn/layers/dense.rs:244-300
def test_dense_identity_8x8():
    y = oracle.dense(np.arange(1, 9, dtype=np.float32)[None], np.eye(8, dtype=np.float16), np.zeros(8, np.float16), relu=True)
    assert y.reshape(-1).tolist() == [1, 2, 3, 4, 5, 6, 7, 8]


def test_dense_1x8_with_offset():
    y = oracle.dense(np.arange(1, 9, dtype=np.float32)[None], np.ones((8, 1), np.float16), np.asarray([-36], np.float16))
    assert y.reshape(-1).tolist() == [0.0]


def test_dense_layout_is_in_major():
    # w[i][o]: a weight matrix that routes input 0 -> output 2 only
    w = np.zeros((3, 4), np.float16)
    w[0, 2] = 1
    y = oracle.dense(np.asarray([[5.0, 7.0, 9.0]], np.float32), w, np.zeros(4, np.float16))
    assert y.reshape(-1).tolist() == [0, 0, 5, 0]


# ---- independent restatement: torch conv2d in fp64, same rounding points
def _torch_forward(t, feats):
    import torch
    import torch.nn.functional as F
    f64 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
    r16 = lambda x: x.to(torch.float16).to(torch.float64)

    def conv(x, name, a1=1.0, a2=0.0, z=None, bias=None):
        w = f64(t[name + ":0"]).permute(0, 3, 1, 2)       # KRSC -> KCRS
        b = f64(t[name + "/offset:0"]) if bias is None else bias
        y = a1 * F.conv2d(x, w, padding=1) + b.view(1, -1, 1, 1)
        if z is not None:
            y = y + a2 * z
        return r16(torch.relu(y))

    nb = 0
    while f"{nb + 2:02d}_residual/conv_1:0" in t:
        nb += 1
    x = f64(feats).view(-1, 19, 19, 32).permute(0, 3, 1, 2)
    a = conv(x, "01_upsample/conv_1")
    for i in range(nb):
        n = f"{i + 2:02d}_residual"
        g = float(np.float32(t[n + "/alpha:0"][0]))
        y = conv(a, n + "/conv_1")
        a = conv(y, n + "/conv_2", a1=g, a2=float(np.float32(1.0) - np.float32(g)), z=a, bias=r16(g * f64(t[n + "/conv_2/offset:0"])))
    h = f"{nb + 2:02d}"
    tau = float(np.float32(1.0) / np.float32(0.709888))
    p1 = conv(a, h + "p_policy/conv_1").permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    p2 = r16(tau * (p1 @ f64(t[h + "p_policy/linear_1:0"])) + r16(tau * f64(t[h + "p_policy/linear_1/offset:0"])))
    pol = torch.softmax(p2, dim=1).to(torch.float16)
    v1 = conv(a, h + "v_value/conv_1").permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    v2 = r16(v1 @ f64(t[h + "v_value/linear_2:0"]) + f64(t[h + "v_value/linear_2/offset:0"]))
    val = torch.tanh(v2).to(torch.float16).view(-1)
    return val.numpy(), pol.numpy(), a.permute(0, 2, 3, 1).reshape(x.shape[0], 361, -1).to(torch.float16).numpy()


def test_whole_network_matches_torch_fp64(small_net):
    from dream_go_b200 import weights
    feats = weights.bernoulli_features(3, seed=11)
    net = oracle.OracleNetwork(small_net)
    assert net.num_blocks == 2
    value, policy, tower = net.forward(feats, want_tower=True)
    tval, tpol, ttower = _torch_forward(small_net, feats)
    # both accumulate in fp64 and round at the same points -> bit-identical
    # except where summation order flips an fp16 rounding (allow a handful)
    mism = np.count_nonzero(tower.view(np.uint16) != ttower.view(np.uint16))
    assert mism <= tower.size * 1e-4, mism
    np.testing.assert_allclose(tower.astype(np.float32), ttower.astype(np.float32), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(policy.astype(np.float32), tpol.astype(np.float32), rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(value.astype(np.float32), tval.astype(np.float32), atol=2e-3)
    assert abs(float(policy.astype(np.float64).sum(1).mean()) - 1.0) < 2e-3


def test_json_roundtrip_feeds_same_network(tmp_path, small_net):
    from dream_go_b200 import weights
    path = str(tmp_path / "dream_go.json")
    weights.dump_json(small_net, path)
    doc = json.load(open(path))
    assert doc["02_residual/alpha:0"]["t"] == "f4" and doc["01_upsample/conv_1:0"]["t"] == "f2"
    loaded = oracle.load_json(path)
    feats = weights.bernoulli_features(1, seed=5)
    v0, p0 = oracle.OracleNetwork(small_net).forward(feats)
    v1, p1 = oracle.OracleNetwork(loaded).forward(feats)
    assert np.array_equal(v0.view(np.uint16), v1.view(np.uint16))
    assert np.array_equal(p0.view(np.uint16), p1.view(np.uint16))
