"""single registration of pair 2 with the search kernel's diagnostics (GOICP_DEBUG=1)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
z = np.load(os.path.join(ROOT, "tests", "golden", "pair2.npz"))
reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.shipped_config(), model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"])
reg.BuildDT(); reg.set_nd(int(z["nd"]))
for _ in range(3):
    t0 = time.perf_counter(); r = reg.Register(); print("Register %.2f ms" % (1e3 * (time.perf_counter() - t0)), r["counters"][:6], flush=True)
