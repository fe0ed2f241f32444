import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
g = ge.load_package()
import importlib
synth = importlib.import_module("goicp_b200.synth")
n, groups = int(sys.argv[1]), int(sys.argv[2])
pairs = synth.bo1_pairs(n, seed=4096)
e = g.Engine()
e.set_batch_options(groups, -1)
res = e.register_batch(g.shipped_config(), pairs)
print("ok", n, groups, [round(r["optError"], 4) for r in res[:6]], e.stats())
