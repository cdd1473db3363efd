"""CPU test of the build-vs-build comparison tool (scripts/dump_networks.py --compare): bit equality, not closeness."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "scripts", "dump_networks.py")


def _run(a, b):
    return subprocess.run([sys.executable, TOOL, "--compare", a, b], capture_output=True, text=True)


def test_compare_is_bitwise(tmp_path):
    x = np.linspace(0, 1, 64, dtype=np.float32)
    m = (x > 0.5).astype(np.uint8)
    a, b, c = (str(tmp_path / n) for n in ("a.npz", "b.npz", "c.npz"))
    np.savez(a, rgb=x, mask=m)
    np.savez(b, rgb=x.copy(), mask=m.copy())
    y = x.copy()
    y[7] = np.nextafter(y[7], np.float32(2.0))          # one ulp: "close" but not identical
    np.savez(c, rgb=y, mask=m)
    same = _run(a, b)
    assert same.returncode == 0 and "all bit-identical" in same.stdout
    diff = _run(a, c)
    assert diff.returncode == 1 and "1 arrays differ" in diff.stdout and "DIFFERENT" in diff.stdout
