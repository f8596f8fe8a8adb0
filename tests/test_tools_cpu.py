"""The soak / fuzz checkers under tools/ in their smallest configuration, so that they keep running (their full-size logs are
under profiles/: r02_soak_*.log, r02_fuzz_*.log)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool,args,expect", [
    ("soak_parity.py", ["--quick"], "ALL IDENTICAL"),
    ("soak_go_parity.py", ["--playouts", "4"], "ALL IDENTICAL"),
    ("fuzz_search_parity.py", ["--cases", "8", "--seed", "9"], "8 cases"),
    ("fuzz_selfplay_config.py", ["--cases", "5", "--seed", "4"], "5 cases, 0 failures"),
])
def test_checker_runs(tool, args, expect):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool)] + args, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    last = out.stdout.strip().splitlines()[-1]
    assert expect in last and "DIFFERENT" not in out.stdout.replace("0 different", ""), last
