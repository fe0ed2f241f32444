"""one single registration of a golden fixture (for ncu launch lists): python scripts/one_register.py pair1|pair2|bunny300|deep_small [relaxed W]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
case = sys.argv[1] if len(sys.argv) > 1 else "bunny300"
W = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cl = {}
if case in ("pair1", "pair2"):
    z = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz")); params = g.shipped_config()
    cl = dict(model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"])
elif case == "bunny300":
    z = np.load(os.path.join(ROOT, "tests", "golden", "bunny.npz")); params = g.upstream_config(distTransSize=300)
else:
    z = np.load(os.path.join(ROOT, "tests", "golden", "deep_small.npz")); params = g.upstream_config(distTransSize=128, MSEThresh=1e-4)
reg = g.GoICP(z["model_xyz"], z["data_xyz"], params, **cl)
reg.BuildDT(); reg.set_nd(int(z["nd"]))
if W: reg.set_search_mode(1, W)
reg.Register()
for _ in range(3):
    t0 = time.perf_counter(); r = reg.Register(); dt = time.perf_counter() - t0
    print(f"  wall {dt*1e3:.2f} ms gpu ms dt {r['gpu_ms_dt']:.2f} bnb {r['gpu_ms_bnb']:.2f} icp {r['gpu_ms_icp']:.2f}")
print(reg.eng.stats() if hasattr(reg, "eng") else "")
print(f"{case} W={W}: Register {dt*1e3:.1f} ms optError {r['optError']:.9g} counters {r['counters'][:6]}")
