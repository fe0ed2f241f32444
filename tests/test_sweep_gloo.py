"""CPU, world_size 2 over gloo: the pair-sharded sweep's partition + result gather (the only N>1 exchange the pair
path has).  The per-rank registration itself is a GPU call and is replaced here by a deterministic stub."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %r)
import __graft_entry__ as ge
ge.load_package()
import importlib
sw = importlib.import_module("goicp_b200.sweep")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
class StubEngine:
    def register_batch(self, params, pairs):
        return [dict(R=np.eye(3) * p["id"], t=np.full(3, p["id"] + 0.5), optError=float(p["id"]) / 4, optComp=p["id"] %% 7, counters=[p["id"]] * 8) for p in pairs]
n = 11
pairs = [dict(id=i) for i in range(n)]
res = sw.sweep(StubEngine(), None, pairs, rank, world)
assert len(res) == n
for i, r in enumerate(res):
    assert r["R"][0, 0] == i and r["t"][1] == i + 0.5 and r["optComp"] == i %% 7 and r["counters"][3] == i and r["optError"] == np.float32(i / 4), (i, r)
lo, hi = sw.shard_range(n, rank, world)
sys.stdout.write("rank %%d block %%d %%d ok\n" %% (rank, lo, hi)); sys.stdout.flush()
dist.destroy_process_group()
"""


def test_shard_range_partition():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.load_package()
    import importlib
    sw = importlib.import_module("goicp_b200.sweep")
    for n in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [sw.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_sweep_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 block 0 6 ok" in out.stdout and "rank 1 block 6 11 ok" in out.stdout


FRONTIER_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %r)
import __graft_entry__ as ge
g = ge.load_package()
import ctypes as C
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# the exchange of the frontier-sharded search (one all-gather of InnerOut records per wave), exercised without a device:
# build the callback exactly as Engine.set_frontier_sharding does, but on a bare object (no goicp handle on a CPU box)
class Fake(g.Engine):
    def __init__(self): self.L = g.lib(); self.h = None
    def check(self, st): assert st in (0, 2), st   # the C call refuses a NULL handle (status 2) -- only the callback is used here
eng = Fake()
eng.set_frontier_sharding(rank, world, None)
nb = 48 * 5
send = (np.arange(nb, dtype=np.uint8) + 7 * rank).astype(np.uint8)
recv = np.zeros(nb * world, np.uint8)
assert eng._ag(send.ctypes.data, recv.ctypes.data, nb, None) == 0
for r in range(world):
    assert np.array_equal(recv[r * nb:(r + 1) * nb], (np.arange(nb, dtype=np.uint8) + 7 * r).astype(np.uint8))
sys.stdout.write("rank %%d exchange ok\n" %% rank); sys.stdout.flush()
dist.destroy_process_group()
"""


def test_frontier_exchange_world2_gloo(tmp_path):
    """the per-wave all-gather of the frontier-sharded registration over gloo, world_size 2"""
    script = tmp_path / "fworker.py"
    script.write_text(FRONTIER_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29733", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 exchange ok" in out.stdout and "rank 1 exchange ok" in out.stdout
