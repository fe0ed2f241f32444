"""debug helper: repeated pair-1 registrations (exact / tree sums) counting wrong optima"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
z = np.load(os.path.join(ROOT, "tests", "golden", "pair1.npz"))
clouds = dict(model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
bad = {0: 0, 1: 0}; tot = {0: 0, 1: 0}
for rep in range(n):
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.shipped_config(), **clouds)
    if rep % 2 == 0:
        reg.BuildDT(); reg.set_nd(int(z["nd"])); reg.Initialize(); reg.ICP(np.eye(3), np.zeros(3))
        reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.shipped_config(), **clouds)
    reg.BuildDT(); reg.set_nd(int(z["nd"]))
    for ex in (1, 0, 0):
        reg.set_options(exact_sums=ex)
        r = reg.Register()
        tot[ex] += 1
        if abs(r["optError"] - 8.453875541687012) > 1e-4:
            bad[ex] += 1
            print("rep", rep, "exact", ex, "optError", r["optError"], "counters", r["counters"][:6], flush=True)
            print("TRACE BAD:\n" + r["trace"], flush=True)
        elif rep == 0 and ex == 0: print("TRACE GOOD:\n" + r["trace"], flush=True)
print("wrong results: exact", bad[1], "/", tot[1], " tree", bad[0], "/", tot[0], flush=True)
