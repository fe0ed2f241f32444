"""where a large-cloud registration spends its GPU time: InnerBnB waves vs ICP"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
import importlib
synth = importlib.import_module("goicp_b200.synth")
def run(name, m, d, nd, params):
    reg = g.GoICP(m, d, params); reg.BuildDT(); reg.set_nd(nd)
    for mode in (0, 1):
        reg.set_search_mode(mode, 256)
        reg.Register()
        t0 = time.perf_counter(); r = reg.Register(); dt = time.perf_counter() - t0
        print(f"{name} {'relaxed' if mode else 'exact'}: Register {dt*1e3:.1f} ms  gpu ms bnb {r['gpu_ms_bnb']:.1f} icp {r['gpu_ms_icp']:.1f}  icp calls {r['counters'][5]} launches {r['counters'][6]} optError {r['optError']:.9g}", flush=True)
z = np.load(os.path.join(ROOT, "tests", "golden", "bunny.npz")); run("bunny300", z["model_xyz"], z["data_xyz"], int(z["nd"]), g.upstream_config(distTransSize=300))
z = np.load(os.path.join(ROOT, "tests", "golden", "deep_small.npz")); run("deep_small", z["model_xyz"], z["data_xyz"], int(z["nd"]), g.upstream_config(distTransSize=128, MSEThresh=1e-4))
p = synth.deep_pair(1236); run("config4", p["model_xyz"], p["data_xyz"], 10000, g.upstream_config(distTransSize=512, MSEThresh=1e-4))
