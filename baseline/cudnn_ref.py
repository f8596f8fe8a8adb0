"""ctypes front-end of baseline/libdg_cudnn_ref.so (cuDNN restatement of dg_nn::forward).
Measurement / test infrastructure only -- never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdg_cudnn_ref.so")
_lib = None


class _View(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_char_p), ("data", C.c_void_p), ("nbytes", C.c_uint64)]


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        # torch ships its own libcudart / libcudnn; make sure the system ones this lib was linked against resolve
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.dgref_last_error.restype = C.c_char_p
        _lib.dgref_info.restype = C.c_char_p
        _lib.dgref_info.argtypes = [C.c_void_p]
        _lib.dgref_create.argtypes = [C.c_int, C.c_int, C.POINTER(_View), C.c_int, C.c_float, C.POINTER(C.c_void_p)]
        _lib.dgref_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.dgref_time_resident.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        _lib.dgref_read_tower.argtypes = [C.c_void_p, C.c_void_p]
        _lib.dgref_destroy.argtypes = [C.c_void_p]
        _lib.dgref_num_blocks.argtypes = [C.c_void_p]
        _lib.dgref_predictor_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(_View), C.c_int, C.c_float, C.POINTER(C.c_void_p)]
        _lib.dgref_predictor_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        _lib.dgref_predictor_destroy.argtypes = [C.c_void_p]
    return _lib


def _views(tensors):
    codes = {np.dtype(np.float16): b"f2", np.dtype(np.float32): b"f4", np.dtype(np.int32): b"i4"}
    keep, views = [], (_View * len(tensors))()
    for i, (name, value) in enumerate(tensors.items()):
        arr = np.ascontiguousarray(value)
        keep.append(arr)
        views[i] = _View(name.encode(), codes[arr.dtype], arr.ctypes.data, arr.nbytes)
    return keep, views


class ReferencePredictor:
    """`NnPredictor` under the reference's own batching rules (predictors/nn.rs:64-107, pool/batch.rs:98-123): batches of at
    most `batch_size` leaves, at most `lanes` = 2 x devices of them in flight, fp16 NHWC features from pageable host memory,
    blocking `nn::forward` on cuDNN.  `.fn` / `.ctx` plug into dream_go_b200.mcts.self_play (a dg_predict_fn)."""

    def __init__(self, tensors, batch_size: int = 16, lanes: int = 2, device: int = 0, temperature: float = 0.709888):
        from dream_go_b200 import mcts
        self._keep, views = _views(tensors)
        self._h = C.c_void_p()
        rc = lib().dgref_predictor_create(device, batch_size, lanes, views, len(tensors), temperature, C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"cuDNN baseline: {lib().dgref_last_error().decode()}")
        self.batch_size, self.lanes = batch_size, lanes
        self.fn = C.cast(lib().dgref_predict, mcts.PREDICT_FN)
        self.ctx = self._h

    def stats(self):
        calls, leaves = C.c_longlong(), C.c_longlong()
        lib().dgref_predictor_stats(self._h, C.byref(calls), C.byref(leaves))
        return {"calls": calls.value, "leaves": leaves.value, "mean_batch": leaves.value / max(calls.value, 1)}

    def close(self):
        if self._h:
            lib().dgref_predictor_destroy(self._h)
            self._h = C.c_void_p()


class CudnnNetwork:
    def __init__(self, tensors, batch: int, device: int = 0, temperature: float = 0.709888):
        self._keep, views = _views(tensors)
        self._h = C.c_void_p()
        self.batch = batch
        rc = lib().dgref_create(device, batch, views, len(tensors), temperature, C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"cuDNN baseline: {lib().dgref_last_error().decode()}")
        self.info = lib().dgref_info(self._h).decode()

    def forward(self, features: np.ndarray):
        f = np.ascontiguousarray(features, dtype=np.float16)
        assert f.size == self.batch * 11552
        value = np.empty((self.batch,), np.float16)
        policy = np.empty((self.batch, 362), np.float16)
        rc = lib().dgref_forward(self._h, f.ctypes.data, value.ctypes.data, policy.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"cuDNN baseline: {lib().dgref_last_error().decode()}")
        return value, policy

    def time_resident(self, iters: int) -> float:
        ms = C.c_float()
        rc = lib().dgref_time_resident(self._h, iters, C.byref(ms))
        if rc != 0:
            raise RuntimeError(f"cuDNN baseline: {lib().dgref_last_error().decode()}")
        return ms.value

    def read_tower(self) -> np.ndarray:
        out = np.empty((self.batch, 361, 128), np.float16)
        lib().dgref_read_tower(self._h, out.ctypes.data)
        return out

    def close(self):
        if self._h:
            lib().dgref_destroy(self._h)
            self._h = C.c_void_p()
