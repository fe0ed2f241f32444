"""batch determinism: the same synthetic sweep registered several times must give identical results (optError, R, t, the
reference-order node counters) for every pair -- speculation and scheduling may differ between runs, results may not"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
import importlib.util
spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "go-icp-protein-cavities_b200", "synth.py")); synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pairs = synth.bo1_pairs(n, seed=4096)
eng = g.Engine(0)
fp = len(sys.argv) > 3 and sys.argv[3] == "fpfh"
params = g.shipped_config(**(dict(cfpfh=1, regularizationFPFH=0.000005) if fp else {}))
eng.batch_upload(params, pairs)
runs = []
for r in range(reps):
    if r == reps - 1: eng.set_options(0, -1, -1)   # last run: tree sums (tolerance)
    res = eng.batch_run()
    runs.append([(x["optError"], tuple(np.asarray(x["R"]).ravel()), tuple(np.asarray(x["t"]).ravel()), tuple(x["counters"][:6])) for x in res])
bad = 0
for k in range(n):
    for r in range(1, reps - 1):
        if runs[r][k] != runs[0][k]:
            bad += 1; print("pair", k, "run", r, runs[r][k][0], runs[0][k][0], runs[r][k][3], runs[0][k][3]); break
tol = sum(1 for k in range(n) if abs(runs[-1][k][0] - runs[0][k][0]) > 1e-4 * max(1.0, runs[0][k][0]))
print("pairs", n, "exact runs", reps - 1, "pairs that differ between exact runs:", bad, "| tree-sum run outside 1e-4:", tol)
