"""Host-side mirror of the reference's `dg_nn` crate surface over the C ABI.

Same names, argument meaning and error behaviour as `src/libdg_nn/lib.rs:33-36`:

* `Network.new()`            -- `network.rs:92-124` (searches the same five paths for `dream_go.json`)
* `Network.get_workspace(n)` -- `network.rs:132-143` (a guard that is returned to the pool on exit)
* `forward(workspace, feats)`-- `graph.rs:123-158`  -> `OutputMap` with `.unwrap() -> (value, policy)`
* `Network.synchronize()`    -- `network.rs:145-159`
* `Error`                    -- `error.rs:19-24`

All arithmetic happens in `libdg_engine.so` (hand-written sm_100a CUDA, `csrc/`).  There is no
CPU path: importing works anywhere, but creating a `Network` without the built library or without
a B200 raises.  numpy is used only to carry host buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DG_ENGINE_LIB") or os.path.join(_HERE, "libdg_engine.so")     # (the override is for A/B runs of two builds)

FEATURE_SIZE = 11552
POLICY_SIZE = 362

DG_OK = 0
FLAG_DEBUG_DIRECT_CONV = 0x1
FLAG_NO_PDL = 0x2
FLAG_LAYERWISE = 0x4
FLAG_NO_ROTATE = 0x8
FLAG_BLOCKING_SYNC = 0x10
FLAG_NO_GRAPH = 0x20
FLAG_TOWER_LATE_A = 0x40
FLAG_SEPARATE_HEAD_CONV = 0x80
LEAF_PRIOR = 0x1


class Error(Exception):
    """`enum Error { CuDNN(Status), Cuda(Error), MalformedWeights, MissingWeights }`"""

    KINDS = {-1: "Cuda", -2: "CuDNN", -3: "MalformedWeights", -4: "MissingWeights", -5: "InvalidArgument"}

    def __init__(self, code: int, message: str = ""):
        self.code = code
        self.kind = self.KINDS.get(code, f"Unknown({code})")
        super().__init__(f"{self.kind}: {message}" if message else self.kind)


class _Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_batch", C.c_int32), ("softmax_temperature", C.c_float),
                ("num_workspaces", C.c_int32), ("flags", C.c_uint32)]


class _TensorView(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_char_p), ("data", C.c_void_p), ("nbytes", C.c_uint64)]


PACKED_DTYPE = np.dtype([("planes", "<u4", (361,)), ("k_bits", "<u2"), ("reserved", "<u2")])   # dg_packed_position

RAW_DTYPE = np.dtype([("black", "<u4", (12,)), ("white", "<u4", (12,)), ("visited", "<u4", (12,)),
                      ("ladder_capture", "<u4", (12,)), ("ladder_escape", "<u4", (12,)), ("hash", "<u8"),
                      ("hash_history", "<u8", (16,)), ("last_move", "<i2", (2,)), ("k_bits", "<u2"), ("to_move", "u1"),
                      ("symmetry", "u1")])   # dg_raw_position, 384 bytes

_lib = None

# every symbol include/dg_engine.h declares: name -> (restype, argtypes)
ABI = {
    "dg_engine_abi_version": (C.c_int32, []),
    "dg_device_count": (C.c_int32, []),
    "dg_current_device": (C.c_int32, []),
    "dg_set_current_device": (C.c_int32, [C.c_int32]),
    "dg_engine_create": (C.c_int32, [C.POINTER(_Config), C.POINTER(C.c_void_p)]),
    "dg_engine_destroy": (None, [C.c_void_p]),
    "dg_engine_load_weights_json": (C.c_int32, [C.c_void_p, C.c_char_p]),
    "dg_engine_load_weights_raw": (C.c_int32, [C.c_void_p, C.POINTER(_TensorView), C.c_int32]),
    "dg_engine_forward_f16": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dg_engine_forward_packed": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dg_engine_forward_raw": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dg_engine_forward_raw_prior": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dg_engine_features_raw": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dg_engine_batch_acquire": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dg_engine_batch_release": (None, [C.c_void_p]),
    "dg_leaf_batch_capacity": (C.c_int32, [C.c_void_p]),
    "dg_leaf_batch_push": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32]),
    "dg_leaf_batch_submit": (C.c_int32, [C.c_void_p, C.c_uint32]),
    "dg_leaf_batch_ready": (C.c_int32, [C.c_void_p]),
    "dg_leaf_batch_wait": (C.c_int32, [C.c_void_p]),
    "dg_leaf_batch_size": (C.c_int32, [C.c_void_p]),
    "dg_leaf_batch_slots": (C.c_void_p, [C.c_void_p]),
    "dg_leaf_batch_value": (C.c_void_p, [C.c_void_p]),
    "dg_leaf_batch_policy": (C.c_void_p, [C.c_void_p]),
    "dg_leaf_batch_legal": (C.c_void_p, [C.c_void_p]),
    "dg_leaf_batch_prior": (C.c_void_p, [C.c_void_p]),
    "dg_leaf_batch_reset": (None, [C.c_void_p]),
    "dg_weights_file_probe": (C.c_int32, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                          C.POINTER(C.c_uint64)]),
    "dg_engine_synchronize": (C.c_int32, [C.c_void_p]),
    "dg_engine_alloc_host": (C.c_void_p, [C.c_void_p, C.c_uint64]),
    "dg_engine_free_host": (None, [C.c_void_p, C.c_void_p]),
    "dg_engine_last_error": (C.c_char_p, [C.c_void_p]),
    "dg_engine_num_blocks": (C.c_int32, [C.c_void_p]),
    "dg_engine_max_batch": (C.c_int32, [C.c_void_p]),
    "dg_engine_num_workspaces": (C.c_int32, [C.c_void_p]),
    "dg_engine_time_resident": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float),
                                            C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "dg_engine_time_e2e": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "dg_engine_debug_read_tower": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "dg_engine_debug_conv_trace": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
    "dg_engine_debug_tower_trace": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
}


def lib() -> C.CDLL:
    """Loads the engine library; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(make -C dream_go_b200/csrc); the engine has no CPU or library fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            if os.environ.get("DG_ENGINE_LIB") and not hasattr(handle, name):
                continue                    # an older build in an A/B run: it lacks the newer entry points
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class OutputMap:
    """`OutputMap<f16>` (`src/libdg_nn/output_map.rs:15-33`)."""

    def __init__(self, value: np.ndarray, policy: np.ndarray):
        self.value, self.policy = value, policy

    def unwrap(self) -> Tuple[np.ndarray, np.ndarray]:
        return self.value, self.policy


class Workspace:
    """`WorkspaceGuard`: exclusive use of one batch size until the `with` block ends.  The engine
    pools the actual device workspaces internally; the guard only pins the batch size, which is all
    the reference's callers rely on (`predictors/nn.rs:93-94`)."""

    def __init__(self, network: "Network", batch_size: int):
        self.network, self.batch_size = network, batch_size

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class Network:
    """Pool of workspaces that can be used for network evaluations (`network.rs:80-160`)."""

    SEARCH_PATHS = ["dream_go.json", "models/dream_go.json", "/usr/share/dreamgo/dream_go.json",
                    "/usr/share/dream_go/dream_go.json"]

    def __init__(self, device: int = 0, max_batch: int = 256, softmax_temperature: float = 0.709888,
                 num_workspaces: int = 2, flags: int = 0):
        self._handle = C.c_void_p()
        cfg = _Config(device, max_batch, softmax_temperature, num_workspaces, flags)
        rc = lib().dg_engine_create(C.byref(cfg), C.byref(self._handle))
        if rc != DG_OK:
            msg = self._last_error()
            if self._handle:
                lib().dg_engine_destroy(self._handle)
                self._handle = C.c_void_p()
            raise Error(rc, msg)
        self.max_batch = max_batch
        self._pinned = []

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def new(cls, **kwargs) -> Optional["Network"]:
        """`Network::new()`: first weights file found in the reference's search order, else None.
        A malformed file raises (the reference panics, `network.rs:111-118`)."""
        exe_json = os.path.splitext(os.path.abspath(sys.argv[0] or "dream_go"))[0] + ".json"
        for path in [exe_json] + cls.SEARCH_PATHS:
            net = cls(**kwargs)
            try:
                net.load_json(path)
                return net
            except Error as err:
                net.close()
                if err.kind != "MissingWeights":
                    raise
        return None

    @classmethod
    def from_tensors(cls, tensors: Dict[str, np.ndarray], **kwargs) -> "Network":
        net = cls(**kwargs)
        net.load_tensors(tensors)
        return net

    def load_json(self, path: str) -> None:
        self._check(lib().dg_engine_load_weights_json(self._handle, os.fsencode(path)))

    def load_tensors(self, tensors: Dict[str, np.ndarray]) -> None:
        codes = {np.dtype(np.float16): b"f2", np.dtype(np.float32): b"f4", np.dtype(np.int32): b"i4", np.dtype(np.int8): b"i1"}
        keep, views = [], (_TensorView * len(tensors))()
        for i, (name, value) in enumerate(tensors.items()):
            arr = np.ascontiguousarray(value)
            keep.append(arr)
            views[i] = _TensorView(name.encode(), codes[arr.dtype], arr.ctypes.data, arr.nbytes)
        self._check(lib().dg_engine_load_weights_raw(self._handle, views, len(tensors)))

    # -- the dg_nn surface ------------------------------------------------------------------------
    def get_workspace(self, batch_size: int) -> Workspace:
        if batch_size < 1 or batch_size > self.max_batch:
            raise Error(-5, f"batch {batch_size} outside 1..{self.max_batch}")
        return Workspace(self, batch_size)

    def synchronize(self) -> None:
        self._check(lib().dg_engine_synchronize(self._handle))

    @property
    def num_blocks(self) -> int:
        return int(lib().dg_engine_num_blocks(self._handle))

    # -- extras used by the harness ---------------------------------------------------------------
    def pinned(self, shape, dtype) -> np.ndarray:
        """numpy view of pinned host memory from dg_engine_alloc_host (freed with the network)."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        ptr = lib().dg_engine_alloc_host(self._handle, nbytes)
        if not ptr:
            raise Error(-1, "dg_engine_alloc_host failed")
        self._pinned.append(ptr)
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def forward_into(self, features: np.ndarray, value: np.ndarray, policy: np.ndarray, packed: bool = False) -> None:
        batch = value.shape[0]
        fn = lib().dg_engine_forward_packed if packed else lib().dg_engine_forward_f16
        self._check(fn(self._handle, features.ctypes.data, batch, value.ctypes.data, policy.ctypes.data))

    def forward_packed(self, positions: np.ndarray) -> OutputMap:
        pos = np.ascontiguousarray(positions, dtype=PACKED_DTYPE).reshape(-1)
        value = np.empty((pos.shape[0],), np.float16)
        policy = np.empty((pos.shape[0], POLICY_SIZE), np.float16)
        self.forward_into(pos, value, policy, packed=True)
        return OutputMap(value, policy)

    def forward_raw(self, positions: np.ndarray):
        """Raw positions -> (OutputMap, legal [n, 361] u8): planes and legal moves are derived on the device."""
        pos = np.ascontiguousarray(positions, dtype=RAW_DTYPE).reshape(-1)
        n = pos.shape[0]
        value = np.empty((n,), np.float16)
        policy = np.empty((n, POLICY_SIZE), np.float16)
        legal = np.empty((n, 361), np.uint8)
        self._check(lib().dg_engine_forward_raw(self._handle, pos.ctypes.data, n, value.ctypes.data, policy.ctypes.data, legal.ctypes.data))
        return OutputMap(value, policy), legal

    def forward_raw_prior(self, positions: np.ndarray):
        """Raw positions -> (OutputMap, legal [n, 361] u8, prior [n, 368] f32): the prior construction runs on the device too."""
        pos = np.ascontiguousarray(positions, dtype=RAW_DTYPE).reshape(-1)
        n = pos.shape[0]
        value = np.empty((n,), np.float16)
        policy = np.empty((n, POLICY_SIZE), np.float16)
        legal = np.empty((n, 361), np.uint8)
        prior = np.empty((n, 368), np.float32)
        self._check(lib().dg_engine_forward_raw_prior(self._handle, pos.ctypes.data, n, value.ctypes.data, policy.ctypes.data,
                                                      legal.ctypes.data, prior.ctypes.data))
        return OutputMap(value, policy), legal, prior

    def features_raw(self, positions: np.ndarray):
        """Feature stage only: (compact planes [n] PACKED_DTYPE, legal [n, 361] u8)."""
        pos = np.ascontiguousarray(positions, dtype=RAW_DTYPE).reshape(-1)
        n = pos.shape[0]
        planes = np.zeros(n, PACKED_DTYPE)
        legal = np.empty((n, 361), np.uint8)
        self._check(lib().dg_engine_features_raw(self._handle, pos.ctypes.data, n, planes.ctypes.data, legal.ctypes.data))
        return planes, legal

    def leaf_batch(self) -> "LeafBatch":
        """One of the engine's in-flight evaluations as a leaf batch (the replacement of `pool::Batcher`)."""
        h = C.c_void_p()
        self._check(lib().dg_engine_batch_acquire(self._handle, C.byref(h)))
        return LeafBatch(self, h)

    def time_resident(self, batch: int, iters: int, tower: bool = True, flush_l2: bool = True) -> Tuple[float, float, int]:
        """(ms for `iters` resident forwards, ms for the residual-conv launches of `iters` forwards, launches/forward)"""
        ms, tms, launches = C.c_float(), C.c_float(), C.c_int32()
        self._check(lib().dg_engine_time_resident(self._handle, batch, iters, int(flush_l2), C.byref(ms),
                                                  C.byref(tms) if tower else None, C.byref(launches)))
        return ms.value, tms.value, launches.value

    def time_e2e(self, features: np.ndarray, steps: int, callers: int = 2) -> float:
        """Wall seconds for `steps` blocking forward_f16 calls issued by `callers` native host threads."""
        feats = np.ascontiguousarray(features, np.float16)
        seconds = C.c_double()
        self._check(lib().dg_engine_time_e2e(self._handle, feats.ctypes.data, feats.shape[0], steps, callers, C.byref(seconds)))
        return seconds.value

    def debug_read_tower(self, layer: int, batch: int) -> np.ndarray:
        out = np.empty((batch, 361, 128), np.float16)
        self._check(lib().dg_engine_debug_read_tower(self._handle, layer, batch, out.ctypes.data))
        return out

    def debug_conv_trace(self, batch: int, ctas: int = 148) -> np.ndarray:
        out = np.zeros((ctas, 3, 64), np.int64)
        self._check(lib().dg_engine_debug_conv_trace(self._handle, batch, out.ctypes.data, out.size))
        return out

    def debug_tower_trace(self, batch: int, ctas: int = 148) -> np.ndarray:
        out = np.zeros((ctas, 3, 1024), np.int64)
        self._check(lib().dg_engine_debug_tower_trace(self._handle, batch, out.ctypes.data, out.size))
        return out

    # -- plumbing -------------------------------------------------------------------------------
    def _last_error(self) -> str:
        if not self._handle:
            return ""
        msg = lib().dg_engine_last_error(self._handle)
        return msg.decode(errors="replace") if msg else ""

    def _check(self, rc: int) -> None:
        if rc != DG_OK:
            raise Error(rc, self._last_error())

    def close(self) -> None:
        if getattr(self, "_handle", None):
            lib().dg_engine_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def probe_weights_file(path: str, name: Optional[str] = None) -> Tuple[int, float, int]:
    """`loader::load` without a device: (number of tensors, scale of `name`, decoded bytes of `name`)."""
    n, scale, nbytes = C.c_int32(), C.c_float(), C.c_uint64()
    rc = lib().dg_weights_file_probe(os.fsencode(path), name.encode() if name else None, C.byref(n), C.byref(scale),
                                     C.byref(nbytes))
    if rc != DG_OK:
        raise Error(rc, path)
    return n.value, scale.value, nbytes.value


def pack_positions(features: np.ndarray) -> np.ndarray:
    """Compact form (`dg_packed_position`) of V1 feature tensors [B, 361, 32] whose planes 2..31 are
    binary and whose planes 0/1 are 0 or the constant k (`features.rs:154-250`)."""
    f = np.ascontiguousarray(features, dtype=np.float16).reshape(-1, 361, 32)
    out = np.zeros((f.shape[0],), dtype=PACKED_DTYPE)
    bits = (f != 0).astype(np.uint32) << np.arange(32, dtype=np.uint32)
    out["planes"] = bits.sum(axis=2, dtype=np.uint32)
    k = f[:, :, :2].reshape(f.shape[0], -1).max(axis=1)
    out["k_bits"] = k.view(np.uint16)
    return out


def forward(workspace: Workspace, features: np.ndarray) -> OutputMap:
    """`dg_nn::forward(&mut Workspace, &[f16]) -> Result<OutputMap<f16>, Error>` (`graph.rs:123-158`).

    `features` holds `batch * 11552` fp16 values, NHWC (`32*(19*y + x) + c`); the result carries
    `value[batch]` (post-tanh) and `policy[batch * 362]` (post-softmax), both fp16."""
    feats = np.ascontiguousarray(features)
    if feats.dtype != np.float16:
        raise Error(-5, "features must be fp16")
    batch = workspace.batch_size
    if feats.size != batch * FEATURE_SIZE:      # debug_assert_eq!, graph.rs:124-125
        raise Error(-5, f"expected {batch * FEATURE_SIZE} features, got {feats.size}")
    value = np.empty((batch,), np.float16)
    policy = np.empty((batch * POLICY_SIZE,), np.float16)
    workspace.network.forward_into(feats, value, policy)
    return OutputMap(value, policy)


def _view(ptr: int, shape, dtype) -> np.ndarray:
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_uint8 * n).from_address(ptr), dtype=dtype).reshape(shape)


class LeafBatch:
    """`dg_leaf_batch`: lock-free pushes of raw positions from any thread, one graph launch per submit, completion read
    from pinned host memory.  Mirrors `Batcher::{push, get_batch}` + `Batch::forward` (pool/batch.rs:61-124)."""

    def __init__(self, network: "Network", handle):
        self.network, self._h = network, handle

    def close(self) -> None:
        if self._h:
            lib().dg_engine_batch_release(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    @property
    def capacity(self) -> int:
        return lib().dg_leaf_batch_capacity(self._h)

    def push(self, positions: np.ndarray) -> int:
        """Returns the index of the first pushed position, or -1 if they do not fit / the batch is sealed."""
        pos = np.ascontiguousarray(positions, dtype=RAW_DTYPE).reshape(-1)
        return int(lib().dg_leaf_batch_push(self._h, pos.ctypes.data, pos.shape[0]))

    def submit(self, prior: bool = False) -> None:
        self.network._check(lib().dg_leaf_batch_submit(self._h, LEAF_PRIOR if prior else 0))

    def ready(self) -> bool:
        return lib().dg_leaf_batch_ready(self._h) == 1

    def wait(self) -> None:
        self.network._check(lib().dg_leaf_batch_wait(self._h))

    def reset(self) -> None:
        lib().dg_leaf_batch_reset(self._h)

    def results(self, prior: bool = False):
        """Copies of (value [n], policy [n, 362], legal [n, 361][, prior [n, 368]]) of the last submit."""
        n = lib().dg_leaf_batch_size(self._h)
        out = [_view(lib().dg_leaf_batch_value(self._h), (n,), np.float16).copy(),
               _view(lib().dg_leaf_batch_policy(self._h), (n, POLICY_SIZE), np.float16).copy(),
               _view(lib().dg_leaf_batch_legal(self._h), (n, 361), np.uint8).copy()]
        if prior:
            out.append(_view(lib().dg_leaf_batch_prior(self._h), (n, 368), np.float32).copy())
        return tuple(out)
