"""One deep registration with the InnerBnB calls of every wave sharded over the ranks of a torchrun job (second shard):
torchrun --nproc-per-node N scripts/frontier_relaxed.py [pair2|bunny|deep_small] [wave_nodes]
wave_nodes > 0: relaxed-order wave mode; 0: the reference's visitation order (wave scheduler with look-ahead calls; for pair2 the
Error*: trace must then equal the reference's).  Every rank applies the same gathered results, so every rank must end with the same
optimum, counters and trace."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import __graft_entry__ as ge
g = ge.load_package()
case = sys.argv[1] if len(sys.argv) > 1 else "pair2"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
z = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
if case == "bunny":
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.upstream_config(distTransSize=300), device=local)
elif case == "deep_small":
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.upstream_config(distTransSize=128, MSEThresh=1e-4), device=local)
else:
    reg = g.GoICP(z["model_xyz"], z["data_xyz"], g.shipped_config(), model_c=z["model_c"], data_c=z["data_c"], model_fpfh=z["model_fpfh"], data_fpfh=z["data_fpfh"], device=local)
if world > 1:
    reg.eng.set_frontier_sharding(rank, world, torch.device("cuda", local))
reg.set_search_mode(1 if W > 0 else 0, W if W > 0 else -1)
reg.BuildDT(); reg.set_nd(int(z["nd"]))
reg.Register()   # warm-up
ts = []
for _ in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); r = reg.Register(); ts.append(time.perf_counter() - t0)
sig = torch.tensor([r["optError"]] + [float(v) for v in r["counters"][:6]], dtype=torch.float64, device=f"cuda:{local}")
same = True
if world > 1:
    ref = sig.clone(); dist.broadcast(ref, 0); same = bool(torch.equal(ref, sig))
evals = r["counters"][2] * int(z["nd"])
trace_ok = None
if W == 0 and case == "pair2":
    trace_ok = g.error_trace(r["trace"]) == list(z["exp_trace"]) and r["optError"] == float(z["exp_optError"]) and r["counters"][:6] == z["exp_counters"][:6].tolist()
print(f"rank {rank}/{world} {case} {'relaxed W=%d' % W if W else 'exact order'}: reference_trace_and_counters {trace_ok} optError {r['optError']:.9g} counters {r['counters'][:6]} same_on_every_rank {same} register {min(ts)*1e3:.1f} ms ({evals / min(ts):.3g} evals/s)", flush=True)
assert same and trace_ok is not False
if world > 1: dist.destroy_process_group()
