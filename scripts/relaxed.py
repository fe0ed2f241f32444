"""exact-order vs relaxed-order (frontier waves) single registrations: wall time, optimum, node counts"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
def cases():
    z = np.load(os.path.join(ROOT, "tests", "golden", "pair1.npz")); cl = dict(model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"])
    yield "pair1", z, g.shipped_config(), cl, float(z["exp_optError"])
    z = np.load(os.path.join(ROOT, "tests", "golden", "pair2.npz")); cl = dict(model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"])
    yield "pair2", z, g.shipped_config(), cl, float(z["exp_optError"])
    z = np.load(os.path.join(ROOT, "tests", "golden", "bunny.npz")); yield "bunny300", z, g.upstream_config(distTransSize=300), {}, float(z["exp300_optError"])
    z = np.load(os.path.join(ROOT, "tests", "golden", "rand.npz")); yield "rand300trim", z, g.upstream_config(distTransSize=300, trimFraction=0.1), {}, float(z["exp300_optError"])
    z = np.load(os.path.join(ROOT, "tests", "golden", "deep_small.npz")); yield "deep_small", z, g.upstream_config(distTransSize=128, MSEThresh=1e-4), {}, float(z["exp128_optError"])
waves = [int(v) for v in sys.argv[1:]] or [64]
for name, z, params, cl, ref in cases():
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], params, **cl)
    reg.BuildDT(); reg.set_nd(int(z["nd"]))
    for mode in [0] + waves:
        reg.set_search_mode(1 if mode else 0, mode if mode else -1)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); r = reg.Register(); ts.append(time.perf_counter() - t0)
        print(f"{name} {'exact order' if not mode else 'relaxed W=%d' % mode}: Register {min(ts)*1e3:.2f} ms optError {r['optError']:.9g} (reference {ref:.9g}) rot pops {r['counters'][3]} calls {r['counters'][0]} subcubes {r['counters'][2]} trace {g.error_trace(r['trace'])[-3:]}", flush=True)
