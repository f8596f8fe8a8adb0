"""Weight-file vocabulary of the dream-go network: tensor names, shapes,
seeded synthetic networks and the `dream_go.json` writer.

The names, dtypes and layouts are the reference's:

* `01_upsample/conv_1:0` .. `NNv_value/linear_2/offset:0`
  (`src/libdg_nn/layers/up_block.rs:42-44`, `residual_block.rs:41-43`,
  `policy_head.rs:49-57`, `value_head.rs:45-53`)
* conv filters are KRSC `[out][3][3][in]` fp16, BatchNorm already folded
  (`contrib/trainer/dream_tf/layers/batch_norm.py:47-74`)
* dense weights are stored `[in][out]` (`contrib/trainer/dream_tf/layers/dense.py:32-39`)
* file = JSON object `{name: {"s": b85(f32 max-abs), "t": "f2"|"f4"|"i4", "v": b85(raw LE bytes)}}`
  (`contrib/trainer/dream_tf/hooks/dump.py:38-67`, `src/libdg_nn/loader.rs:36-100`)

No trained network ships with the reference, so every test and benchmark in
this repository runs on seeded synthetic weights in this exact format.
"""
from __future__ import annotations

import base64
import json
from typing import Dict

import numpy as np

NUM_FEATURES = 32          # src/libdg_go/utils/features.rs:88-90
NUM_POINTS = 361
POLICY_SIZE = 362
FEATURE_SIZE = NUM_POINTS * NUM_FEATURES   # 11,552 fp16 per position
DEFAULT_CHANNELS = 128     # src/libdg_nn/layers/common.rs:22
DEFAULT_SAMPLES = 8        # src/libdg_nn/layers/common.rs:25
VALUE_SAMPLES = 2          # src/libdg_nn/layers/value_head.rs:42
SOFTMAX_TEMPERATURE = 0.709888   # src/libdg_utils/config.rs:176-177


def tensor_shapes(num_blocks: int = 9, channels: int = DEFAULT_CHANNELS,
                  samples: int = DEFAULT_SAMPLES) -> Dict[str, tuple]:
    """Name -> shape of every fp16 tensor `Builder::get_workspace` looks up
    (`src/libdg_nn/graph.rs:50-96`)."""
    shapes = {
        "01_upsample/conv_1:0": (channels, 3, 3, NUM_FEATURES),
        "01_upsample/conv_1/offset:0": (channels,),
    }
    for i in range(num_blocks):
        n = f"{i + 2:02d}_residual"
        shapes[f"{n}/conv_1:0"] = (channels, 3, 3, channels)
        shapes[f"{n}/conv_1/offset:0"] = (channels,)
        shapes[f"{n}/conv_2:0"] = (channels, 3, 3, channels)
        shapes[f"{n}/conv_2/offset:0"] = (channels,)
    h = f"{num_blocks + 2:02d}"
    shapes[f"{h}p_policy/conv_1:0"] = (samples, 3, 3, channels)
    shapes[f"{h}p_policy/conv_1/offset:0"] = (samples,)
    shapes[f"{h}p_policy/linear_1:0"] = (NUM_POINTS * samples, POLICY_SIZE)
    shapes[f"{h}p_policy/linear_1/offset:0"] = (POLICY_SIZE,)
    shapes[f"{h}v_value/conv_1:0"] = (VALUE_SAMPLES, 3, 3, channels)
    shapes[f"{h}v_value/conv_1/offset:0"] = (VALUE_SAMPLES,)
    shapes[f"{h}v_value/linear_2:0"] = (NUM_POINTS * VALUE_SAMPLES, 1)
    shapes[f"{h}v_value/linear_2/offset:0"] = (1,)
    return shapes


def synthetic_network(seed: int = 20261017, num_blocks: int = 9, channels: int = DEFAULT_CHANNELS,
                      samples: int = DEFAULT_SAMPLES, gate: float | str = 0.5,
                      head_scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded random network with the reference tensor names.

    Conv filters ~ N(0, 2/(9*Cin)) (He), offsets ~ N(0, 0.1), as SURVEY.md
    section 8d config 2 prescribes; `gate` is 0.5 (the reference default,
    `residual_block.rs:43,50`) or "random" for g ~ U[0,1] per block.
    """
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in tensor_shapes(num_blocks, channels, samples).items():
        if name.endswith("/offset:0"):
            w = rng.normal(0.0, 0.1, size=shape)
        elif "linear" in name:
            w = rng.normal(0.0, head_scale * np.sqrt(1.0 / shape[0]), size=shape)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            w = rng.normal(0.0, np.sqrt(2.0 / fan_in), size=shape)
        out[name] = w.astype(np.float16)
    for i in range(num_blocks):
        g = float(rng.uniform(0.0, 1.0)) if gate == "random" else float(gate)
        out[f"{i + 2:02d}_residual/alpha:0"] = np.asarray([g], dtype=np.float32)
    out["num_blocks:0"] = np.asarray([num_blocks], dtype=np.int32)
    out["num_channels:0"] = np.asarray([channels], dtype=np.int32)
    out["num_samples:0"] = np.asarray([samples], dtype=np.int32)
    return out


_TYPE_CODES = {np.dtype(np.float16): "f2", np.dtype(np.float32): "f4",
               np.dtype(np.int32): "i4", np.dtype(np.int8): "i1"}


def dump_json(tensors: Dict[str, np.ndarray], path: str, model_name: str | None = "synthetic") -> None:
    """Writes `tensors` in the trainer's dump format
    (`contrib/trainer/dream_tf/hooks/dump.py:38-67`): `pad=True` base85 of the
    little-endian bytes, `"s"` = fp32 max-abs (stored by the engine, never used
    by `forward`)."""
    doc = {}
    for name, value in tensors.items():
        value = np.ascontiguousarray(value)
        code = _TYPE_CODES[value.dtype]
        max_abs = np.asarray(np.max(np.abs(value.astype(np.float64))) if value.size else 0.0, dtype="<f4")
        doc[name] = {
            "s": base64.b85encode(max_abs.tobytes(), pad=True).decode("ascii"),
            "t": code,
            "v": base64.b85encode(value.astype(value.dtype.newbyteorder("<")).tobytes(), pad=True).decode("ascii"),
        }
    if model_name is not None:
        doc["model_name:0"] = model_name   # plain JSON string, ignored by the loader (loader.rs:47)
    with open(path, "w") as fh:
        json.dump(doc, fh, sort_keys=True)


def bernoulli_features(batch: int, seed: int = 1, p: float = 0.2) -> np.ndarray:
    """The reference's own forward-bench input distribution: iid Bernoulli(0.2)
    in {0,1} over all 11,552 fp16 features (`src/dg_tests/benches/batch_sizes.rs:44-49`)."""
    rng = np.random.default_rng(seed)
    return (rng.random((batch, NUM_POINTS, NUM_FEATURES)) < p).astype(np.float16)
