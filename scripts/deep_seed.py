"""config #4 (synthetic deep search): which seeds leave the rotation queue non-trivial (the initial ICP from the identity does not
reach the optimum)?  Prints per seed the counters and the pose error against the generating motion.
    python scripts/deep_seed.py <nm> <nd> <S> <seed0> <nseeds> [mse]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
import importlib
synth = importlib.import_module("goicp_b200.synth")
nm, nd, S, seed0, n = [int(v) for v in sys.argv[1:6]]
mse = float(sys.argv[6]) if len(sys.argv) > 6 else 1e-4
for seed in range(seed0, seed0 + n):
    p = synth.deep_pair(seed, nm=nm, nd=nd)
    reg = g.GoICP(p["model_xyz"], p["data_xyz"], g.upstream_config(distTransSize=S, MSEThresh=mse))
    t0 = time.perf_counter(); reg.BuildDT(); tdt = time.perf_counter() - t0
    t0 = time.perf_counter(); r = reg.Register(); treg = time.perf_counter() - t0
    dR = np.abs(r["R"] - p["R_true"]).max(); dt = np.abs(r["t"] - p["t_true"]).max()
    print(f"seed {seed}: DT {tdt*1e3:.1f} ms Register {treg*1e3:.1f} ms optError {r['optError']:.6g} counters {r['counters'][:6]} |R-Rtrue| {dR:.3g} |t-ttrue| {dt:.3g} trace {g.error_trace(r['trace'])}", flush=True)
