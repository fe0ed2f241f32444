import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__ as ge
g = ge.load_package()
from oracle import pyoracle as po
from conftest import rand_rot
for nd in (1500, 3000):
    rng = np.random.default_rng(21)
    model = rng.normal(size=(500, 3)); model = (0.7 * model / np.abs(model).max()).astype(np.float32)
    data = rng.normal(size=(nd, 3)); data = (0.6 * data / np.abs(data).max()).astype(np.float32)
    kw = dict(distTransSize=32, trimFraction=0.2)
    reg = g.GoICP(model, data, g.upstream_config(**kw))
    o = po.Oracle("port", model, data, po.upstream_config(**kw))
    reg.BuildDT(); o.build_dt(); reg.set_nd(nd); o.set_nd(nd); reg.Initialize(); o.initialize()
    for kind in ("port", "ref"):
        if not po.available(kind): continue
        oo = po.Oracle(kind, model, data, po.upstream_config(**kw)); oo.build_dt(); oo.set_nd(nd); oo.initialize()
        e, R, t, corr = reg.ICP(np.eye(3), np.zeros(3))
        eo, Ro, to, co = oo.icp(np.eye(3), np.zeros(3))
        print(nd, kind, "err", e, eo, "dR", np.abs(R - Ro).max(), "dt", np.abs(t - to).max(), "corr equal", np.array_equal(corr, co), int((np.asarray(corr) != np.asarray(co)).sum()))
