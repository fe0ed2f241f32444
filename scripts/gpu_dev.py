"""Development check on a GPU box: CUDA path vs the CPU oracle, verbose.  (tests/ holds the formal parity tests.)"""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
from oracle import pyoracle as po

def G(n): return np.load(os.path.join(ROOT, "tests", "golden", n + ".npz"))
def rand_rot(rng):
    v = rng.uniform(-np.pi, np.pi, 3).astype(np.float32)
    while np.linalg.norm(v) > np.pi: v = rng.uniform(-np.pi, np.pi, 3).astype(np.float32)
    t = np.linalg.norm(v); k = v / t
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return (np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * K @ K).astype(np.float32)

def section(name, fn):
    t = time.time()
    try:
        fn(); print(f"[ok] {name}  ({time.time()-t:.2f}s)", flush=True)
    except Exception:
        print(f"[FAIL] {name}", flush=True); traceback.print_exc(); sys.stdout.flush()

def pair_case(name, fp=False):
    z = G(name)
    clouds = dict(model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"])
    kw = dict(cfpfh=1, regularizationFPFH=0.000005) if fp else {}
    return z, clouds, g.shipped_config(**kw), po.shipped_config(**kw)

def t_components(name="pair1", fp=False):
    z, clouds, gp, op = pair_case(name, fp)
    nd = int(z["nd"])
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], gp, **clouds)
    o = po.Oracle("port", z["model_xyz"], z["data_xyz"], op, **clouds)
    info = reg.BuildDT(); o.build_dt(); oi = o.dt_info()
    assert info.scale == oi["scale"] and info.xMin == oi["xMin"] and info.zMax == oi["zMax"], (info.scale, oi)
    d, near, cc = reg.dt_download(); od, ooff, onear, occ = o.dt_download()
    print("  dt dist equal:", (d == od).all(), " nearest equal:", (near == onear).all(), "mismatch", int((near != onear).any(1).sum()), " cellc equal:", (cc == occ).all())
    assert (d == od).all() and (near == onear).all() and (cc == occ).all()
    assert (d == z["exp_dt_dist"]).all() and (near == z["exp_dt_near"]).all()
    # separable builder
    reg.set_options(use_dt_replay=0); reg.BuildDT(); d2, near2, _ = reg.dt_download()
    print("  separable: dist mismatches", int((d2 != od).sum()), "max abs", float(np.abs(d2 - od).max()), " nearest mismatches", int((near2 != onear).any(1).sum()))
    reg.set_options(use_dt_replay=1); reg.BuildDT()
    rng = np.random.default_rng(0)
    q = rng.uniform(-1.5, 1.5, (5000, 3))
    a, ca = reg.Distance(q); b, cb = o.dt_distance(q)
    assert (a == b).all() and (ca == cb).all()
    reg.set_nd(nd); o.set_nd(nd); reg.Initialize(); o.initialize()
    assert (reg.weights() == o.weights()).all(), np.abs(reg.weights() - o.weights()).max()
    assert (reg.maxRotDis() == o.maxrotdis()).all()
    assert reg.thresholds() == (o.ssethresh(), o.inliernum())
    for it in range(4):
        R = rand_rot(rng); level = [-1, 0, 2, 5][it]
        w = np.float32(2.0 ** -rng.integers(0, 5)); tc = np.concatenate([rng.uniform(-0.5, 0.5 - w, (200, 3)), np.full((200, 1), w)], 1).astype(np.float32)
        ub, lb, inc, fpm = reg.eval_bounds(R, level, tc); oub, olb, oinc, ofp = o.eval_leaf(R, level, tc)
        rel = max(np.abs(ub - oub).max() / max(oub.max(), 1e-9), np.abs(lb - olb).max() / max(olb.max(), 1e-9))
        print(f"  eval_bounds level {level}: rel err {rel:.2e} incomp equal {(inc == oinc).all()} fpfh equal {(fpm == ofp).all()}")
        assert rel < 1e-5 and (inc == oinc).all() and (fpm == ofp).all()
    Rs = np.stack([rand_rot(rng) for _ in range(12)]); lv = np.array([-1, 1] * 6, np.int32); oe = np.full(12, 25.0, np.float32)
    for exact in (1, 0):
        reg.set_options(exact_sums=exact)
        err, tn, ps = reg.InnerBnB(Rs, lv, oe)
        cnt0 = None
        for k in range(12):
            e, t4 = o.inner_bnb(Rs[k], int(lv[k]), 25.0)
            if exact: assert err[k] == np.float32(e) and (lv[k] >= 0 or err[k] == 25.0 or (tn[k] == t4).all()), (k, err[k], e, tn[k], t4)
            else: assert abs(err[k] - e) <= 1e-5 * max(abs(e), 1), (k, err[k], e)
        print(f"  inner_bnb exact={exact}: ok  pops {ps[:,0].tolist()}")
    reg.set_options(exact_sums=1)
    e, R, t, corr = reg.ICP(np.eye(3), np.zeros(3)); eo, Ro, to, co = o.icp(np.eye(3), np.zeros(3))
    print("  icp err", e, eo, "dR", np.abs(R - Ro).max(), "dt", np.abs(t - to).max(), "corr equal", (corr == co).all())
    assert e == eo and np.abs(R - Ro).max() < 1e-12 and (corr == co).all()

def t_register(name, fp=False, exact=1, spec=32):
    z, clouds, gp, op = pair_case(name, fp)
    pre = "expf_" if fp else "exp_"
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], gp, **clouds); reg.set_options(exact_sums=exact, spec_width=spec)
    reg.BuildDT(); reg.set_nd(int(z["nd"]))
    t = time.time(); r = reg.Register(); dt = time.time() - t
    print(f"  {name} fp={fp} exact={exact}: optError {r['optError']} (ref {z[pre+'optError']}) comp {r['optComp']} ({z[pre+'optComp']}) counters {r['counters']} ref {z[pre+'counters'][:6].tolist()} wall {dt:.3f}s timings {reg.eng.timings()}")
    print("   trace", g.error_trace(r["trace"]), list(z[pre + "trace"]))
    assert abs(r["optError"] - float(z[pre + "optError"])) <= 1e-5 * float(z[pre + "optError"])
    assert np.abs(r["R"] - z[pre + "R"]).max() < 1e-5 and np.abs(r["t"] - z[pre + "t"]).max() < 1e-5
    assert r["optComp"] == int(z[pre + "optComp"])
    if exact: assert r["counters"][:6] == z[pre + "counters"][:6].tolist()

def t_demo(name, S, exact=1, upload_ref_dt=True):
    z = G(name); nd = int(z["nd"]); trim = float(z["trim"])
    gp = g.upstream_config(trimFraction=trim, distTransSize=S); op = po.upstream_config(trimFraction=trim, distTransSize=S)
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], gp); reg.set_options(exact_sums=exact)
    t = time.time(); reg.BuildDT(); tdt = time.time() - t
    if upload_ref_dt:
        o = po.Oracle("port", z["model_xyz"], z["data_xyz"], op); t = time.time(); o.build_dt(); tod = time.time() - t
        od, _, onear, _ = o.dt_download(); d, near, _ = reg.dt_download()
        nm = int((d != od).sum())
        print(f"  {name} S={S}: GPU separable DT {tdt:.3f}s (oracle 8SED {tod:.2f}s) dist mismatches {nm}/{S**3} max abs {float(np.abs(d-od).max()):.3g}; nearest mismatches {int((near != onear).any(1).sum())}")
        reg.dt_upload(od, onear)
    reg.set_nd(nd)
    t = time.time(); r = reg.Register(); dt = time.time() - t
    pre = f"exp{S}_"
    print(f"  {name} S={S} exact={exact} refDT={upload_ref_dt}: optError {r['optError']} (ref {z[pre+'optError']}) counters {r['counters']} ref {z[pre+'counters'][:6].tolist()} wall {dt:.3f}s {reg.eng.timings()}")
    print("   trace", g.error_trace(r["trace"]), list(z[pre + "trace"]))
    assert abs(r["optError"] - float(z[pre + "optError"])) <= 2e-5 * float(z[pre + "optError"])
    assert np.abs(r["R"] - z[pre + "R"]).max() < 1e-4 and np.abs(r["t"] - z[pre + "t"]).max() < 1e-4

def t_xform():
    z = G("pair1"); e = g.Engine()
    cen, mean, mx = e.normalizeMolCloud(z["src_raw"])
    assert np.abs(mean - z["src_mean"]).max() < 1e-12 and abs(mx - float(z["src_maxnorm"])) < 1e-12, (mean, z["src_mean"], mx, z["src_maxnorm"])
    sc = e.scaleCloud(cen, float(z["scale"])); assert np.abs(sc - z["src_scaled"]).max() < 1e-15
    tt = e.rescaleCloud(float(z["scale"]), z["tgt_mean"], z["src_mean"], z["exp_R"], z["exp_t"])
    print("  rescaled t", tt, z["exp_rescaled_t"]); assert np.abs(tt - z["exp_rescaled_t"]).max() < 1e-3
    rot = e.applyTransformationProtein(z["protein_xyz"], z["exp_R"], tt)
    rm = e.computeRMSD(z["aligned_xyz"][:len(z["rot_xyz"])], z["rot_xyz"]) if False else None
    print("  rot max diff vs shipped rot file", np.abs(rot - z["rot_xyz"]).max())

which = sys.argv[1:] or ["comp", "reg", "demo", "xform"]
if "comp" in which:
    section("components pair1", lambda: t_components("pair1"))
    section("components pair1 fpfh", lambda: t_components("pair1", True))
if "reg" in which:
    section("register pair1", lambda: t_register("pair1"))
    section("register pair1 fast", lambda: t_register("pair1", exact=0))
    section("register pair1 fpfh", lambda: t_register("pair1", True))
    section("register pair2", lambda: t_register("pair2"))
    section("register pair2 fast", lambda: t_register("pair2", exact=0))
if "demo" in which:
    section("rand S=64", lambda: t_demo("rand", 64))
    section("bunny S=100 refDT", lambda: t_demo("bunny", 100))
    section("bunny S=100 own DT", lambda: t_demo("bunny", 100, upload_ref_dt=False))
    section("bunny S=100 fast", lambda: t_demo("bunny", 100, exact=0))
if "xform" in which:
    section("transformation", t_xform)

def t_bunny300():
    z = G("bunny"); nd = int(z["nd"])
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.upstream_config(distTransSize=300))
    for exact in (1, 0):
        reg.set_options(exact_sums=exact)
        t = time.time(); reg.BuildDT(); tdt = time.time() - t; reg.set_nd(nd)
        t = time.time(); r = reg.Register(); dt = time.time() - t
        print(f"  bunny S=300 exact={exact}: DT {tdt:.3f}s register {dt:.3f}s optError {r['optError']:.9g} (ref {float(z['exp300_optError']):.9g}) counters {r['counters']} ref {z['exp300_counters'][:6].tolist()} {reg.eng.timings()} {reg.eng.stats()}")
        print("   trace", g.error_trace(r["trace"]), list(z["exp300_trace"]))
        print("   dR", np.abs(r["R"] - z["exp300_R"]).max(), "dt", np.abs(r["t"] - z["exp300_t"]).max())

def t_deep():
    import importlib
    synth = importlib.import_module("goicp_b200.synth")
    p = synth.deep_pair()
    reg = g.GoICP(p["model_xyz"], p["data_xyz"], g.upstream_config(distTransSize=512, MSEThresh=1e-4))
    reg.set_options(exact_sums=int(os.environ.get("DEEP_EXACT", "0")))
    t = time.time(); reg.BuildDT(); tdt = time.time() - t
    t = time.time(); r = reg.Register(); dt = time.time() - t
    Rt, tt = p["R_true"], p["t_true"]
    print(f"  deep: DT 512^3 {tdt:.3f}s register {dt:.3f}s optError {r['optError']:.6g} counters {r['counters']} {reg.eng.timings()} {reg.eng.stats()}")
    print("   |R - R_true|", np.abs(r["R"] - Rt).max(), "|t - t_true|", np.abs(r["t"] - tt).max(), g.error_trace(r["trace"]))

if "bunny300" in which: section("bunny S=300", t_bunny300)
if "deep" in which: section("deep config #4", t_deep)
