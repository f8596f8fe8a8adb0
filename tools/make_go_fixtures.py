"""Generates tests/golden/go_fixtures.npz from the reference's DATA (run in the build container, where
/root/reference exists; the result is committed because the GPU box has no /root/reference):

* zobrist   [3][420] u64  -- the random constants of src/libdg_go/zobrist.rs:18 (data, needed only to pin
                             the oracle against the hash KATs of dg_tests/tests/real_games.rs:49,74,117)
* games     the 99 game records of src/dg_tests/fixtures/example_games.sgf as flat (colour, 19*y+x) arrays,
            y flipped exactly like dg_tests/tests/common/mod.rs:54-64 (`Point::new(x, 18 - y)`)
* kat_*     the three games quoted in dg_tests/tests/real_games.rs with their expected final hashes

    python tools/make_go_fixtures.py
"""
import os
import re
import numpy as np

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "go_fixtures.npz")
MOVE = re.compile(r";([BW])\[([a-z]*)\]")


def parse_moves(src):
    colors, moves = [], []
    for m in MOVE.finditer(src):
        c = 1 if m.group(1) == "B" else 2
        s = m.group(2)
        x = ord(s[0]) - 97 if len(s) > 0 else 19
        y = ord(s[1]) - 97 if len(s) > 1 else 19
        colors.append(c)
        moves.append(19 * (18 - y) + x if x < 19 and y < 19 else 361)
    return np.array(colors, np.uint8), np.array(moves, np.uint16)


def main():
    text = open(f"{REF}/libdg_go/zobrist.rs").read()
    vals = [int(v, 16) for v in re.findall(r"0x([0-9a-fA-F]{16})", text)]
    assert len(vals) == 3 * 420, len(vals)
    out = {"zobrist": np.array(vals, np.uint64).reshape(3, 420)}

    colors, moves, offsets, komi = [], [], [0], []
    for line in open(f"{REF}/dg_tests/fixtures/example_games.sgf", encoding="utf-8", errors="replace"):
        if not line.strip():
            continue
        c, m = parse_moves(line)
        colors.append(c)
        moves.append(m)
        offsets.append(offsets[-1] + len(m))
        km = re.search(r"KM\[([^\]]*)\]", line)
        k = float(km.group(1)) if km else 7.5
        komi.append(k / 100.0 if abs(k) > 100 else k)     # the file writes 6.5 as "650"
    out["games_colors"] = np.concatenate(colors)
    out["games_moves"] = np.concatenate(moves)
    out["games_offsets"] = np.array(offsets, np.int64)
    out["games_komi"] = np.array(komi, np.float32)

    src = open(f"{REF}/dg_tests/tests/real_games.rs").read()
    games = re.findall(r'playout_game\(r#"(.*?)"#, None\);\s*assert_eq!\(board\.zobrist_hash\(\), (\d+)', src, re.S)
    assert len(games) == 3
    for i, (sgf, want) in enumerate(games):
        c, m = parse_moves(sgf)
        out[f"kat{i}_colors"], out[f"kat{i}_moves"] = c, m
        out[f"kat{i}_hash"] = np.array([int(want)], np.uint64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(komi), "games,", offsets[-1], "moves")


if __name__ == "__main__":
    main()
