"""One deep registration with its rotation frontier sharded over the ranks of a torchrun job (second shard).
   torchrun --nproc-per-node N scripts/frontier_shard.py [pair2|bunny]
Every rank must end with the reference's optimum, node counters and trace (the search is identical to the 1-GPU one)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import __graft_entry__ as ge
g = ge.load_package()
case = sys.argv[1] if len(sys.argv) > 1 else "pair2"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
z = np.load(os.path.join(ROOT, "tests", "golden", ("bunny" if case == "bunny" else case) + ".npz"))
if case == "bunny":
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.upstream_config(distTransSize=300), device=local); pre = "exp300_"
else:
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.shipped_config(), model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"], device=local); pre = "exp_"
if world > 1:
    reg.eng.set_frontier_sharding(rank, world, torch.device("cuda", local))
reg.BuildDT(); reg.set_nd(int(z["nd"]))
reg.Register()   # warm-up
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter(); r = reg.Register(); dt = time.perf_counter() - t0
ok = r["optError"] == float(z[pre + "optError"]) and r["counters"][:6] == z[pre + "counters"][:6].tolist() and g.error_trace(r["trace"]) == list(z[pre + "trace"])
print(f"rank {rank}/{world} {case}: optError {r['optError']:.9g} counters {r['counters'][:6]} identical_to_reference {ok} register {dt*1e3:.1f} ms gpu_ms_bnb {r['gpu_ms_bnb']:.1f}", flush=True)
assert ok
if world > 1: dist.destroy_process_group()
