"""CPU: the C-ABI library loads and exports exactly the symbols include/goicp_b200.h declares; struct layouts match the
oracle's; the product fails loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "goicp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(goicp_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(g):
    syms = _header_symbols()
    assert sorted(g.ABI_SYMBOLS) == syms
    L = g.lib()
    for s in syms:
        assert hasattr(L, s), s
    out = subprocess.check_output(["nm", "-D", "--defined-only", g.LIB_PATH]).decode()
    exported = set(re.findall(r" T (goicp_[a-z0-9_]+)", out))
    assert exported == set(syms), exported ^ set(syms)


def test_struct_layouts(g, po):
    assert C.sizeof(g.Params) == C.sizeof(po.Params) == 80
    assert [f[0] for f in g.Params._fields_] == [f[0] for f in po.Params._fields_]
    assert C.sizeof(g.Result) == 9 * 8 + 3 * 8 + 4 + 4 + 8 * 8 + 16 + 12 + 4
    p, q = g.shipped_config(), po.shipped_config()
    assert bytes(p) == bytes(q)
    assert bytes(g.upstream_config()) == bytes(po.upstream_config())


def test_no_cpu_fallback(g):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g.GoICPError) as e:
        g.Engine()
    assert e.value.status == 1 and "no CPU fallback" in str(e.value)


def test_product_never_touches_oracle():
    """the product path may not import, link, load or name anything under oracle/ (it is test infrastructure)"""
    pkg = os.path.join(ROOT, "go-icp-protein-cavities_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                assert "oracle" not in open(os.path.join(dp, f)).read().lower(), (dp, f)


def test_cpp_dropin_builds(g):
    """the header-only C++ mirror of the reference classes compiles and links against the C-ABI library"""
    subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "examples")])
    assert os.path.exists(os.path.join(ROOT, "examples", "goicp_demo")) and os.path.exists(os.path.join(ROOT, "examples", "GoICP_b200"))


def test_bench_reference_arm_contract(po):
    """`bench.py --impl reference` (the reference's own CPU implementation on the host cores: no GPU involved, so it runs here):
    one JSON line with the keys of the bench contract, the same metric / unit / config as the GPU arm, `impl: reference`, a
    `cpu_baseline` describing the run and an `e2e` block repeating the line's value with zero copy bytes."""
    import json, sys
    if not (po.available("ref") or po.available("port")):
        pytest.skip("no oracle library built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-pairs", "4"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cube_point_bound_evals_per_sec" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("BO1-shaped") and d["config"]["pairs_per_gpu"] == 4096
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "pairs" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
