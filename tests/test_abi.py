"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/stnerf.h declares,
and fails loudly (no fallback) without a device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from stnerf_b200 import _lib as L
    if not os.path.isfile(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return L


def test_header_symbols_are_exported():
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "stnerf.h")).read()
    declared = set(re.findall(r"\b(stnerf_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libstnerf_b200.so does not export %s" % name
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)


def test_struct_layouts_match_header():
    L = _lib()
    # sizes implied by include/stnerf.h with STNERF_MAX_LAYERS = 8
    assert ctypes.sizeof(L.ModelDesc) == 4 + 4 * 8 + 4 + 4
    assert ctypes.sizeof(L.Scene) == (96 + 96) + 32 + 32 + 96 + 32 + 32 + 32 + 12 + 5 * 4 + 4 + 4
    hdr = open(os.path.join(ROOT, "include", "stnerf.h")).read()
    assert "#define STNERF_MAX_LAYERS %d" % L.MAX_LAYERS in hdr
    assert "#define STNERF_MAX_N1 %d" % L.MAX_N1 in hdr
    assert "#define STNERF_MAX_S %d" % L.MAX_S in hdr


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    L = _lib()
    lib = L.lib()
    h = ctypes.c_void_p()
    d = L.ModelDesc()
    d.n_layers = 3
    assert lib.stnerf_create(ctypes.byref(h), ctypes.byref(d)) == -2            # STNERF_ENODEVICE
    assert b"CUDA device" in lib.stnerf_strerror(-2)
    from stnerf_b200 import NativeRenderer, StnerfError, ops
    with pytest.raises(StnerfError):
        NativeRenderer(3, [False, True, True])
    with pytest.raises(StnerfError):
        ops.positional_encoding(torch.zeros(4, 3), 4)                            # CPU tensor -> loud failure


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under st-nerf_b200/ may reference it."""
    pkg = os.path.join(ROOT, "st-nerf_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|stnerf_oracle|reference_shim|oracle\.)")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                # the only legitimate mention is the data directory oracle/_ref/ckpt (checkpoint copies, not code)
                assert not bad.search(src), os.path.join(dp, f)
    # the native arm of bench.py, the examples and the scripts stay clear of it too (the CPU leg of bench.py is the exception)
    for rel in ("examples/taekwondo_demo_b200.py", "scripts/profile_render.py", "scripts/bench_stages.py"):
        assert not bad.search(open(os.path.join(ROOT, rel)).read()), rel
    # bench.py: the oracle / reference harness (and the test-side `cases` module, which imports it) may only be imported inside
    # the functions of the CPU leg
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    cpu_leg = {"cpu_sample", "cpu_reference_rate", "cpu_leg_parity"}

    def visit(node, fn):
        for ch in ast.iter_child_nodes(node):
            name = ch.name if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef)) else fn
            if isinstance(ch, ast.ImportFrom) and (ch.module or "").split(".")[0] in ("oracle",):
                assert fn in cpu_leg, (fn, ch.module)
            if isinstance(ch, ast.Import) and any(a.name.split(".")[0] in ("oracle", "cases") for a in ch.names):
                assert fn in cpu_leg, (fn, [a.name for a in ch.names])
            visit(ch, name)

    visit(tree, None)


def test_precision_enum_matches_python_map():
    """include/stnerf.h STNERF_PREC_* values == the names the Python side accepts (cfg.MODEL.B200_PRECISION / --precision)."""
    import re
    from stnerf_b200 import _lib as L
    hdr = open(os.path.join(ROOT, "include", "stnerf.h")).read()
    enum = {m.group(1): int(m.group(2)) for m in re.finditer(r"STNERF_PREC_(\w+)\s*=\s*(\d+)", hdr)}
    assert enum == {"FP32_SIMT": 0, "TC_3XF16": 1, "TC_F16": 2, "TC_MIXED": 3, "TC_3XF16_CF": 4}
    assert L.PRECISIONS["fp32"] == enum["FP32_SIMT"] and L.PRECISIONS["exact"] == enum["TC_3XF16"]
    assert L.PRECISIONS["fast"] == enum["TC_F16"] and L.PRECISIONS["mixed"] == enum["TC_MIXED"]
    assert L.PRECISIONS["exact_cf"] == enum["TC_3XF16_CF"]
