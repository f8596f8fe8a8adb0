"""ctypes front-end of the CPU oracle (`oracle/libdg_oracle.so`).

TEST INFRASTRUCTURE ONLY -- imported by tests/, by bench.py's cpu_baseline /
`--impl reference` leg and by `__graft_entry__.smoke()`, never by the product
package `dream_go_b200`.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from typing import Dict, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdg_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in os.listdir(_HERE) if f.endswith((".c", ".cpp", ".h"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libdg_oracle.so"])
    return _LIB_PATH


class _Net(C.Structure):
    _fields_ = [
        ("num_blocks", C.c_int32), ("channels", C.c_int32), ("features", C.c_int32),
        ("policy_samples", C.c_int32), ("value_samples", C.c_int32), ("tau", C.c_float),
        ("up_w", C.c_void_p), ("up_b", C.c_void_p),
        ("res_w1", C.POINTER(C.c_void_p)), ("res_b1", C.POINTER(C.c_void_p)),
        ("res_w2", C.POINTER(C.c_void_p)), ("res_b2", C.POINTER(C.c_void_p)),
        ("res_gate", C.POINTER(C.c_float)),
        ("pol_conv_w", C.c_void_p), ("pol_conv_b", C.c_void_p), ("pol_fc_w", C.c_void_p), ("pol_fc_b", C.c_void_p),
        ("val_conv_w", C.c_void_p), ("val_conv_b", C.c_void_p), ("val_fc_w", C.c_void_p), ("val_fc_b", C.c_void_p),
    ]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.dg_oracle_f32_to_f16.restype = C.c_uint16
        _lib.dg_oracle_f32_to_f16.argtypes = [C.c_float]
        _lib.dg_oracle_f16_to_f32.restype = C.c_float
        _lib.dg_oracle_f16_to_f32.argtypes = [C.c_uint16]
        _lib.dg_oracle_b85_decode.restype = C.c_long
        _lib.dg_oracle_b85_decode.argtypes = [C.c_char_p, C.c_long, C.c_void_p, C.c_long]
        _lib.dg_oracle_b85_encode.restype = C.c_long
        _lib.dg_oracle_b85_encode.argtypes = [C.c_void_p, C.c_long, C.c_char_p, C.c_long]
        _lib.dg_oracle_conv3x3.restype = None
        _lib.dg_oracle_conv3x3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        _lib.dg_oracle_dense.restype = None
        _lib.dg_oracle_dense.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_float, C.c_int, C.c_void_p]
        _lib.dg_oracle_forward.restype = C.c_int
        _lib.dg_oracle_forward.argtypes = [C.POINTER(_Net), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        _lib.dg_oracle_num_threads.restype = C.c_int
    return _lib


# ---------------------------------------------------------------- scalars / b85

def f32_to_f16_bits(x: float) -> int:
    return int(lib().dg_oracle_f32_to_f16(C.c_float(x)))


def f16_bits_to_f32(bits: int) -> float:
    return float(lib().dg_oracle_f16_to_f32(C.c_uint16(bits)))


def b85_decode(text: bytes) -> bytes:
    buf = C.create_string_buffer(len(text) // 5 * 4 + 4)
    n = lib().dg_oracle_b85_decode(text, len(text), buf, len(buf))
    if n < 0:
        raise ValueError("invalid base85 input")
    return buf.raw[:n]


def b85_encode(raw: bytes) -> bytes:
    buf = C.create_string_buffer(len(raw) // 4 * 5 + 5)
    n = lib().dg_oracle_b85_encode(raw, len(raw), buf, len(buf))
    if n < 0:
        raise ValueError("length must be a multiple of 4")
    return buf.raw[:n]


_DTYPES = {"f2": "<f2", "f4": "<f4", "i4": "<i4", "i1": "i1"}


def load_json(path: str) -> Dict[str, np.ndarray]:
    """Oracle-side reader of `dream_go.json`, following `src/libdg_nn/loader.rs:36-100`:
    plain-string entries are ignored, `"t"` picks the element type, `"v"` is
    base85 of little-endian elements.  Returns flat arrays (possibly carrying
    one padding element -- "size by the descriptor, not the decoded length")."""
    with open(path, "r") as fh:
        doc = json.load(fh)
    if not doc:
        raise ValueError("MissingWeights")
    out = {}
    for name, entry in doc.items():
        if isinstance(entry, str):
            continue
        if set(entry) - {"s", "t", "v"} or entry.get("t") not in _DTYPES:
            raise ValueError("MalformedWeights")
        out[name] = np.frombuffer(b85_decode(entry["v"].encode("ascii")), dtype=_DTYPES[entry["t"]]).copy()
    return out


# ---------------------------------------------------------------- layers

def conv3x3(x: np.ndarray, w_krsc: np.ndarray, bias: np.ndarray, a1: float = 1.0, a2: float = 0.0,
            z: np.ndarray | None = None, relu: bool = True) -> np.ndarray:
    """x: [N, H, H, Cin] (fp16-representable), w: [Cout,3,3,Cin] fp16, bias [Cout] fp16 -> [N,H,H,Cout] float32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, wh, _, cin = x.shape
    w = np.ascontiguousarray(w_krsc, dtype=np.float16)
    cout = w.shape[0]
    b = np.ascontiguousarray(bias, dtype=np.float16)
    zz = None if z is None else np.ascontiguousarray(z, dtype=np.float32)
    y = np.empty((n, wh, wh, cout), dtype=np.float32)
    lib().dg_oracle_conv3x3(x.ctypes.data, n, wh, cin, w.ctypes.data, b.ctypes.data, cout,
                            a1, a2, None if zz is None else zz.ctypes.data, int(relu), y.ctypes.data)
    return y


def dense(x: np.ndarray, w_in_out: np.ndarray, bias: np.ndarray, a1: float = 1.0, relu: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, n_in = x.shape
    w = np.ascontiguousarray(w_in_out, dtype=np.float16).reshape(n_in, -1)
    n_out = w.shape[1]
    b = np.ascontiguousarray(bias, dtype=np.float16)
    y = np.empty((n, n_out), dtype=np.float32)
    lib().dg_oracle_dense(x.ctypes.data, n, n_in, w.ctypes.data, b.ctypes.data, n_out, a1, int(relu), y.ctypes.data)
    return y


# ---------------------------------------------------------------- whole network

class OracleNetwork:
    """Holds a tensor dict (reference names) and evaluates `dg_nn::forward` on the CPU."""

    def __init__(self, tensors: Dict[str, np.ndarray], softmax_temperature: float = 0.709888):
        self._keep = []
        t = tensors

        def f16(name, count=None):
            a = np.ascontiguousarray(t[name]).reshape(-1)
            if a.dtype != np.float16:
                raise ValueError(f"{name}: expected f2")
            if count is not None:
                if a.size < count:
                    raise ValueError(f"{name}: too short")
                a = a[:count]          # drop base85 padding
            a = np.ascontiguousarray(a)
            self._keep.append(a)
            return a.ctypes.data

        ch = int(np.asarray(t["num_channels:0"]).reshape(-1)[0]) if "num_channels:0" in t else 128
        sm = int(np.asarray(t["num_samples:0"]).reshape(-1)[0]) if "num_samples:0" in t else 8
        nb = 0
        while f"{nb + 2:02d}_residual/conv_1:0" in t and f"{nb + 2:02d}_residual/conv_2:0" in t:
            nb += 1                    # graph.rs:76-96 -- stop at the first missing block
        self.num_blocks, self.channels, self.samples = nb, ch, sm
        net = _Net()
        net.num_blocks, net.channels, net.features = nb, ch, 32
        net.policy_samples, net.value_samples = sm, 2
        net.tau = float(np.float32(1.0) / np.float32(softmax_temperature))   # f32 division, policy_head.rs:46
        net.up_w = f16("01_upsample/conv_1:0", ch * 9 * 32)
        net.up_b = f16("01_upsample/conv_1/offset:0", ch)
        arr = lambda: (C.c_void_p * max(nb, 1))()
        w1, b1, w2, b2 = arr(), arr(), arr(), arr()
        gates = (C.c_float * max(nb, 1))()
        for i in range(nb):
            n = f"{i + 2:02d}_residual"
            w1[i] = f16(f"{n}/conv_1:0", ch * 9 * ch)
            b1[i] = f16(f"{n}/conv_1/offset:0", ch)
            w2[i] = f16(f"{n}/conv_2:0", ch * 9 * ch)
            b2[i] = f16(f"{n}/conv_2/offset:0", ch)
            gates[i] = float(np.asarray(t[f"{n}/alpha:0"], dtype=np.float32).reshape(-1)[0]) if f"{n}/alpha:0" in t else 0.5
        self._keep += [w1, b1, w2, b2, gates]
        net.res_w1, net.res_b1, net.res_w2, net.res_b2 = w1, b1, w2, b2
        net.res_gate = gates
        h = f"{nb + 2:02d}"
        net.pol_conv_w = f16(f"{h}p_policy/conv_1:0", sm * 9 * ch)
        net.pol_conv_b = f16(f"{h}p_policy/conv_1/offset:0", sm)
        net.pol_fc_w = f16(f"{h}p_policy/linear_1:0", 361 * sm * 362)
        net.pol_fc_b = f16(f"{h}p_policy/linear_1/offset:0", 362)
        net.val_conv_w = f16(f"{h}v_value/conv_1:0", 2 * 9 * ch)
        net.val_conv_b = f16(f"{h}v_value/conv_1/offset:0", 2)
        net.val_fc_w = f16(f"{h}v_value/linear_2:0", 722)
        net.val_fc_b = f16(f"{h}v_value/linear_2/offset:0", 1)
        self._net = net

    def forward(self, features: np.ndarray, want_tower: bool = False, want_blocks: bool = False):
        """features: [B,361,32] fp16 -> (value fp16 [B], policy fp16 [B,362][, tower fp16 [B,361,C]][, blocks])."""
        f = np.ascontiguousarray(features, dtype=np.float16).reshape(-1, 361, 32)
        b = f.shape[0]
        value = np.empty((b,), dtype=np.float16)
        policy = np.empty((b, 362), dtype=np.float16)
        tower = np.empty((b, 361, self.channels), dtype=np.float16) if want_tower else None
        blocks = np.empty((self.num_blocks + 1, b, 361, self.channels), dtype=np.float16) if want_blocks else None
        rc = lib().dg_oracle_forward(C.byref(self._net), f.ctypes.data, b, value.ctypes.data, policy.ctypes.data,
                                     None if tower is None else tower.ctypes.data,
                                     None if blocks is None else blocks.ctypes.data)
        if rc != 0:
            raise MemoryError("oracle forward failed")
        out: Tuple = (value, policy)
        if want_tower:
            out += (tower,)
        if want_blocks:
            out += (blocks,)
        return out


def num_threads() -> int:
    return int(lib().dg_oracle_num_threads())
