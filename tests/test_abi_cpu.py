"""CPU-side checks of the boundary: the library loads, exports every symbol the header declares,
and the device-independent weight-file reader follows the reference loader (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from dream_go_b200 import nn, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(header="dg_engine.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    handle = ctypes.CDLL(nn.LIB_PATH)
    names = header_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/dg_engine.h but not exported"
    assert set(names) == set(nn.ABI), "python binding table and header disagree"


def test_library_exports_every_go_symbol():
    from dream_go_b200 import go
    handle = ctypes.CDLL(nn.LIB_PATH)
    names = header_functions("dg_go.h")
    assert len(names) >= 24
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/dg_go.h but not exported"
    assert set(names) == set(go.ABI), "python binding table and header disagree"


def test_library_exports_every_mcts_symbol():
    from dream_go_b200 import mcts
    handle = ctypes.CDLL(nn.LIB_PATH)
    names = [n for n in header_functions("dg_mcts.h") if n not in ("dg_predict_fn", "dg_predict_raw_fn", "dg_predict_prior_fn")]
    assert len(names) >= 11
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/dg_mcts.h but not exported"
    assert set(names) == set(mcts.ABI), "python binding table and header disagree"


def test_abi_version():
    assert nn.lib().dg_engine_abi_version() == 1


def test_device_helpers_without_gpu():
    """`Device::all()` of the shim: no usable device here, and asking never throws."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = nn.lib()
    assert L.dg_device_count() == 0
    assert L.dg_current_device() == -1
    assert L.dg_set_current_device(0) == -1      # DG_ERR_CUDA


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nn.Error) as err:
        nn.Network(max_batch=4)
    assert err.value.kind == "Cuda"


def test_loader_kat(tmp_path):
    # src/libdg_nn/loader.rs:131-141
    path = tmp_path / "net.json"
    path.write_text('{"11v_value/linear_2/offset:0": {"s": "(^d>V", "t": "f2", "v": "(^d>V"}}')
    count, scale, nbytes = nn.probe_weights_file(str(path), "11v_value/linear_2/offset:0")
    assert count == 1
    assert np.float32(scale) == np.float32(0.13704996)
    assert nbytes == 4


def test_loader_errors(tmp_path):
    # empty input is an error (loader.rs:124-129); a missing file is MissingWeights (loader.rs:110-116)
    empty = tmp_path / "empty.json"
    empty.write_text("")
    with pytest.raises(nn.Error) as err:
        nn.probe_weights_file(str(empty))
    assert err.value.kind == "MissingWeights"
    with pytest.raises(nn.Error) as err:
        nn.probe_weights_file(str(tmp_path / "nope.json"))
    assert err.value.kind == "MissingWeights"
    for bad in ['{"a": {"t": "f8", "v": "(^d>V"}}',        # unknown type
                '{"a": {"t": "f2", "v": "(^d> "}}',        # invalid base85 character
                '{"a": {"t": "f2", "q": "(^d>V"}}',        # unknown attribute
                '{"a": [1, 2]}', '[]']:
        p = tmp_path / "bad.json"
        p.write_text(bad)
        with pytest.raises(nn.Error) as err:
            nn.probe_weights_file(str(p))
        assert err.value.kind == "MalformedWeights", bad
    only_strings = tmp_path / "s.json"
    only_strings.write_text('{"model_name:0": "x"}')
    with pytest.raises(nn.Error) as err:
        nn.probe_weights_file(str(only_strings))
    assert err.value.kind == "MissingWeights"


def test_dumped_network_is_readable(tmp_path, small_net):
    path = tmp_path / "dream_go.json"
    weights.dump_json(small_net, str(path))
    count, scale, nbytes = nn.probe_weights_file(str(path), "01_upsample/conv_1:0")
    assert count == len(small_net)                       # model_name:0 is a plain string and is skipped
    assert nbytes == 128 * 9 * 32 * 2
    assert np.float32(scale) == np.float32(np.abs(small_net["01_upsample/conv_1:0"].astype(np.float64)).max())
    _, _, nb = nn.probe_weights_file(str(path), "04p_policy/linear_1/offset:0")
    assert nb == 362 * 2                                 # 724 bytes: already a multiple of 4
    _, _, nb = nn.probe_weights_file(str(path), "04v_value/linear_2/offset:0")
    assert nb == 4                                       # one fp16 + base85 padding


def test_pack_positions_roundtrip():
    rng = np.random.default_rng(3)
    f = (rng.random((5, 361, 32)) < 0.3).astype(np.float16)
    k = np.float16(0.9667)
    f[:, :, 0] = np.where(np.arange(5)[:, None] % 2 == 0, k, 0)
    f[:, :, 1] = np.where(np.arange(5)[:, None] % 2 == 1, k, 0)
    p = nn.pack_positions(f)
    assert p.dtype.itemsize == 1448
    back = ((p["planes"][:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.float16)
    back[:, :, :2] *= p["k_bits"].view(np.float16)[:, None, None]
    assert np.array_equal(back, f)


def test_forward_argument_checks_need_no_device():
    ws = nn.Workspace(network=None, batch_size=2)
    with pytest.raises(nn.Error):
        nn.forward(ws, np.zeros((2 * 11552 - 1,), np.float16))
    with pytest.raises(nn.Error):
        nn.forward(ws, np.zeros((2 * 11552,), np.float32))


def test_search_and_selfplay_reject_bad_arguments():
    import ctypes as C
    from dream_go_b200 import go, mcts
    L = mcts.lib()
    board = go.Board(7.5)
    opt = mcts._SearchOptions(0, 1, 10, 1, 0.25, 0.8, 1, None, None, 0, -1.0, None)
    null_fn = C.cast(None, mcts.PREDICT_FN)
    assert L.dg_mcts_predict(null_fn, None, C.byref(opt), None, board._h, 1, None, None, None, None) == -5
    cfg = mcts._SelfPlayConfig(0, 4, 10, 1, 10, 1, 0, 0, 0.25, 0.8, 1, 0.0, 0, 0)          # num_games = 0
    stats = mcts._SelfPlayStats()
    rnd = mcts.RandomPredictor()
    assert L.dg_selfplay_run(rnd.fn, None, C.byref(cfg), C.byref(stats), None, 0) == -5
    # a failing predictor's status comes back unchanged
    failing = mcts.PREDICT_FN(lambda ctx, pos, n, v, p: -2)
    assert L.dg_mcts_predict(failing, None, C.byref(opt), None, board._h, 1, None, None, None, None) == -2
    cfg = mcts._SelfPlayConfig(2, 2, 10, 1, 10, 1, 0, 0, 0.25, 0.8, 1, 0.0, 0, 0)
    assert L.dg_selfplay_run(failing, None, C.byref(cfg), C.byref(stats), None, 0) == -2


def test_engine_entry_points_reject_null_engine():
    L = nn.lib()
    assert L.dg_engine_forward_f16(None, None, 1, None, None) == -5
    assert L.dg_engine_forward_raw(None, None, 1, None, None, None) == -5
    assert L.dg_engine_batch_acquire(None, None) == -5
    assert L.dg_leaf_batch_push(None, None, 1) == -5 and L.dg_leaf_batch_submit(None, 0) == -5
    assert L.dg_leaf_batch_ready(None) == -5 and L.dg_leaf_batch_wait(None) == -5 and L.dg_leaf_batch_size(None) == 0
    assert L.dg_engine_max_batch(None) == 0 and L.dg_engine_num_workspaces(None) == 0
    from dream_go_b200 import mcts
    assert mcts.lib().dg_selfplay_run_engine(None, 0, 0, None, None, None, 0) == -5


def test_bench_self_play_summary_on_recorded_lines():
    """bench.py's compact self-play summary (the part of the line the driver's record keeps) on the lines recorded in profiles/."""
    import json
    import sys
    sys.path.insert(0, ROOT)
    import bench
    for name, has_ref in (("r02_bench_n1.json", True), ("r02_bench_n8.json", False)):
        line = json.load(open(os.path.join(ROOT, "profiles", name)))
        out = bench.self_play_summary(line["self_play"])
        assert out["configs2_32_games_per_gpu"][0] > 0 and out["games128_per_gpu"][1] > out["configs2_32_games_per_gpu"][1]
        assert ("cudnn_reference_32_games_batch16" in out) == has_ref
        if has_ref:
            assert out["configs2_vs_cudnn_reference_batch16"] > 5
        assert len(json.dumps(out)) < 900


# ---- struct layouts: the headers as gcc lays them out, the ctypes mirrors, and the `repr(C)` mirrors of INTEGRATION.md ---------

RUST_TYPES = {"i32": (4, 4), "u32": (4, 4), "f32": (4, 4), "i64": (8, 8), "u64": (8, 8), "f64": (8, 8), "u16": (2, 2), "u8": (1, 1), "i16": (2, 2)}


def rust_layout(struct_name: str):
    """Field offsets and size of a `#[repr(C)] pub struct` of INTEGRATION.md under the C layout rules."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    body = re.search(r"pub struct " + struct_name + r"\s*\{(.*?)\}", text, re.S).group(1)
    body = re.sub(r"//[^\n]*", "", body)
    at, align_max, offsets = 0, 1, {}
    for name, ty in re.findall(r"(\w+)\s*:\s*([^,]+?)\s*(?:,|$)", body.replace("\n", " ")):
        m = re.fullmatch(r"\[(\w+);\s*(\d+)\]", ty)
        if ty.startswith("*"):
            size, align = 8, 8
        elif m:
            size, align = RUST_TYPES[m.group(1)][0] * int(m.group(2)), RUST_TYPES[m.group(1)][1]
        else:
            size, align = RUST_TYPES[ty]
        at = (at + align - 1) // align * align
        offsets[name] = at
        at += size
        align_max = max(align_max, align)
    return offsets, (at + align_max - 1) // align_max * align_max


def c_layouts(tmp_path):
    """sizeof / offsetof of the ABI structs, printed by a C program compiled against include/*.h."""
    structs = {
        "dg_search_options": ["search", "deterministic", "num_rollout", "probes_per_round", "dirichlet_noise", "temperature", "seed", "noise",
                              "leaf_symmetries", "n_leaf_symmetries", "choose_at", "cache", "device_ladders"],
        "dg_selfplay_config": ["num_games", "num_parallel", "num_rollout", "probes_per_round", "max_plies", "num_threads", "ex_it",
                               "num_ex_it_rollout", "dirichlet_noise", "temperature", "seed", "max_seconds", "cache_capacity", "num_groups",
                               "cache_shared"],
        "dg_selfplay_stats": ["games_finished", "moves", "evals", "rounds", "searches", "seconds", "eval_seconds", "mean_batch", "digest",
                              "cache_hits"],
        "dg_packed_position": ["planes", "k_bits", "reserved"],
        "dg_raw_position": ["black", "white", "visited", "ladder_capture", "ladder_escape", "hash", "hash_history", "last_move", "k_bits",
                            "to_move", "symmetry"],
    }
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "dg_mcts.h"', "int main(void) {"]
    for s, fields in structs.items():
        src.append(f'  printf("{s} size %zu\\n", sizeof({s}));')
        for f in fields:
            src.append(f'  printf("{s} {f} %zu\\n", offsetof({s}, {f}));')
    src += ["  return 0;", "}"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    import subprocess
    subprocess.check_call(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = {}
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        s, f, v = line.split()
        out.setdefault(s, {})[f] = int(v)
    return out


def test_struct_layouts_of_headers_ctypes_and_rust_mirrors_agree(tmp_path):
    from dream_go_b200 import mcts
    C = c_layouts(tmp_path)
    for cname, ctype, rust in (("dg_search_options", mcts._SearchOptions, "DgSearchOptions"),
                               ("dg_selfplay_config", mcts._SelfPlayConfig, "DgSelfPlayConfig"),
                               ("dg_selfplay_stats", mcts._SelfPlayStats, "DgSelfPlayStats")):
        want = dict(C[cname])
        size = want.pop("size")
        assert ctypes.sizeof(ctype) == size, cname
        assert {n: getattr(ctype, n).offset for n, _ in ctype._fields_} == want, cname
        offsets, rsize = rust_layout(rust)
        assert offsets == want and rsize == size, (cname, offsets, want)
    # the two position formats that cross the boundary by value
    assert C["dg_packed_position"]["size"] == nn.PACKED_DTYPE.itemsize == 361 * 4 + 4
    offsets, rsize = rust_layout("DgPackedPosition")
    assert rsize == C["dg_packed_position"]["size"] and offsets["k_bits"] == C["dg_packed_position"]["k_bits"]
    assert C["dg_raw_position"]["size"] == 384 and C["dg_raw_position"]["hash"] % 8 == 0
    offsets, rsize = rust_layout("DgRawPosition")
    want = dict(C["dg_raw_position"])
    assert rsize == want.pop("size") and offsets == want
    assert nn.RAW_DTYPE.itemsize == 384 and {n: nn.RAW_DTYPE.fields[n][1] for n in nn.RAW_DTYPE.names if n in want} == \
        {n: want[n] for n in nn.RAW_DTYPE.names if n in want}


INVALID_ARGUMENT = -5                       # DG_ERR_INVALID_ARGUMENT (include/dg_engine.h)


def test_search_entry_points_reject_what_the_reference_types_cannot_express():
    """`Color`, `SearchOptions` and `Point` are types in the reference; at the C boundary they are integers and pointers."""
    import ctypes as C
    from dream_go_b200 import go as pgo, mcts as pm
    L = pm.lib()
    board = pgo.Board(7.5)
    fn = C.cast(L.dg_random_predict, pm.PREDICT_FN)
    v, i, t, e = C.c_float(), C.c_int32(), C.c_void_p(), C.c_int64()

    def search(predictor=fn, options=None, board_handle=board._h, color=1, **fields):
        opt = pm._SearchOptions(0, 1, 12, 2, 0.25, 0.8, 1, None, None, 0, -1.0, None, 0)
        for name, value in fields.items():
            setattr(opt, name, value)
        t.value = None
        rc = L.dg_mcts_predict(predictor, None, C.byref(opt) if options is None else options, None, board_handle, color,
                               C.byref(v), C.byref(i), C.byref(t), C.byref(e))
        if rc == 0:
            L.dg_tree_free(t)
        return rc

    assert search() == 0 and 0 <= i.value <= 361
    assert search(predictor=C.cast(None, pm.PREDICT_FN)) == INVALID_ARGUMENT
    assert search(board_handle=None) == INVALID_ARGUMENT
    for color in (0, 3, -1, 77):
        assert search(color=color) == INVALID_ARGUMENT
    assert search(search=2) == INVALID_ARGUMENT and search(search=-1) == INVALID_ARGUMENT
    assert search(n_leaf_symmetries=3) == INVALID_ARGUMENT          # a count without the array
    assert search(probes_per_round=0) == 0 and search(num_rollout=0) == 0 and search(probes_per_round=-4, num_rollout=-9) == 0
    # trees: indices outside 0..361
    _, _, tree, _ = pm.predict(pm.RandomPredictor(), board, 1, deterministic=True, num_rollout=20, seed=3)
    before = tree.children()[0].copy()
    for index in (-1, 362, 1 << 20):
        L.dg_tree_disqualify(tree._h, index)
    assert (tree.children()[0] == before).all()
    assert L.dg_tree_forward(tree.release(), 4000) is None                      # consumed, no sub-tree
    assert L.dg_tree_forward(None, 3) is None
    # self-play configurations that cannot run
    st = pm._SelfPlayStats()
    for games, parallel in ((0, 1), (1, 0), (-3, 2)):
        cfg = pm._SelfPlayConfig(games, parallel, 10, 1, 5, 1, 0, 10, 0.25, 0.8, 1, 0.0, 0, 0, 0)
        assert L.dg_selfplay_run(fn, None, C.byref(cfg), C.byref(st), None, 0) == INVALID_ARGUMENT
    cfg = pm._SelfPlayConfig(1, 1, 10, 1, 5, 1, 0, 10, 0.25, 0.8, 1, 0.0, 0, 0, 0)
    assert L.dg_selfplay_run(C.cast(None, pm.PREDICT_FN), None, C.byref(cfg), C.byref(st), None, 0) == INVALID_ARGUMENT
    assert L.dg_selfplay_run_engine(None, 1, 0, C.byref(cfg), C.byref(st), None, 0) == INVALID_ARGUMENT
    assert L.dg_selfplay_run(fn, None, None, C.byref(st), None, 0) == INVALID_ARGUMENT
