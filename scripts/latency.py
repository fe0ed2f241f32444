"""wall time of single registrations (warm), per case"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
def run(name, fp=False, S=None, reps=5, **opt):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    if name in ("pair1", "pair2"):
        kw = dict(cfpfh=1, regularizationFPFH=0.000005) if fp else {}
        reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.shipped_config(**kw), model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"]); pre = "expf_" if fp else "exp_"
    else:
        reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.upstream_config(distTransSize=S, trimFraction=float(z["trim"]))); pre = f"exp{S}_"
    reg.set_options(**opt)
    t0 = time.perf_counter(); reg.BuildDT(); tdt = time.perf_counter() - t0
    reg.set_nd(int(z["nd"]))
    walls = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = reg.Register(); walls.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); reg.BuildDT(); tdt2 = time.perf_counter() - t0; reg.set_nd(int(z["nd"]))
    same = r["counters"][:6] == z[pre + "counters"][:6].tolist()
    print(f"{name} fp={fp} S={S} {opt}: DT {tdt2*1e3:.2f} ms (first {tdt*1e3:.1f}) Register min {min(walls)*1e3:.2f} ms med {sorted(walls)[len(walls)//2]*1e3:.2f} ms optError {r['optError']:.9g} ref {float(z[pre+'optError']):.9g} counters_identical {same}", flush=True)
run("pair1"); run("pair1", fp=True); run("pair2"); run("rand", S=64); run("bunny", S=100); run("bunny", S=300); run("bunny", S=300, exact_sums=0)
