#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json on N B200s of one node.

Workload (config.workload): a sweep of synthetic cavity pairs shaped like the BO1 dataset (BASELINE.json configs[4]
shape, per-pair settings of configs[1]: the shipped config.txt with its incompatibility term, Nd = all source points,
DT 20^3), `--pairs` pairs per GPU per step (weak scaling: pairs are sharded across ranks, no data-path collective).
One step = BuildDT + Initialize + Register of every pair of the rank's block.

  value     cube.point bound evals / s (1 eval = jly_goicp.cpp:369-381 once), whole job, inputs resident in HBM
  e2e       the same through the C-ABI call that takes HOST buffers (goicp_register_batch): host preprocessing,
            host->device copies and the result read-back inside the timed region
  roofline  dominant kernel = inner_bnb_kernel; achieved = evals x 4 B (one DT voxel per eval) / its
            CUDA-event time, against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, compiled from its sources) on a bounded sample
            of the same pairs, one process per host core

`--impl reference` times that CPU arm alone on the same config/metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per worker stream (default 8 aliases them)
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC, UNIT = "cube_point_bound_evals_per_sec", "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=4096, help="cavity pairs per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--exact", type=int, default=1, help="1: reference-identical sequential float sums; 0: warp-tree sums")
    ap.add_argument("--fpfh", type=int, default=0, help="1: add the c-FPFH term (cfpfh=1, regularizationFPFH=5e-6)")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs in the CPU-baseline sample (0: two per core, at least 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=4096)
    ap.add_argument("--groups", type=int, default=-1, help="worker streams per GPU (-1: library default)")
    ap.add_argument("--slots", type=int, default=-1, help="pairs advanced in lock-step per stream (-1: library default)")
    ap.add_argument("--spec", type=int, default=-1, help="speculation width in rotation nodes (-1: library default)")
    return ap.parse_args()


def config_dict(args, extra=None):
    c = {"workload": "BO1-shaped synthetic cavity pairs (configs[4] shape; per-pair settings of configs[1]: shipped config.txt, "
                     "Nd=all, DT 20^3%s)" % (", c-FPFH term" if args.fpfh else ""),
         "pairs_per_gpu": args.pairs, "sharding": "pairs across ranks, no collective on the data path",
         "sums": "exact (reference order)" if args.exact else "warp-tree",
         "l2": "flushed between timed steps (256 MiB write)"}
    if extra:
        c.update(extra)
    return c


# ---- CPU arm: the reference compiled from its own sources, one process per core ------------------------------------
def _cpu_worker(job):
    kind, pair, fpfh = job
    from oracle import pyoracle as po
    kw = dict(cfpfh=1, regularizationFPFH=0.000005) if fpfh else {}
    o = po.Oracle(kind, pair["model_xyz"], pair["data_xyz"], po.shipped_config(**kw), model_c=pair["model_c"], data_c=pair["data_c"],
                  model_fpfh=pair["model_fpfh"], data_fpfh=pair["data_fpfh"])
    r = o.register(pair["nd"])
    o.close()
    return r["counters"][2] * pair["nd"], r["optError"]


def cpu_arm(pairs, fpfh, cores):
    """returns (evals/s, seconds, kind, optErrors)"""
    import multiprocessing as mp
    from oracle import pyoracle as po
    kind = "ref" if po.available("ref") else "port"
    po._lib(kind)
    jobs = [(kind, p, fpfh) for p in pairs]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, jobs[:cores])   # touch pages / load libs
        t0 = time.perf_counter()
        out = pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    evals = sum(o[0] for o in out)
    return evals / dt, dt, ("reference" if kind == "ref" else "port"), [o[1] for o in out]


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons, "samples": len(self.rows)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import __graft_entry__ as ge
    g = ge.load_package()
    import importlib
    synth = importlib.import_module("goicp_b200.synth")
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        n = args.cpu_pairs or max(8, 2 * cores)
        pairs = synth.bo1_pairs(n, seed=args.seed)
        vals = []
        for _ in range(args.warmup):
            cpu_arm(pairs[:cores], args.fpfh, cores)
        t_all = 0.0
        for _ in range(args.steps):
            v, dt, kind, _ = cpu_arm(pairs, args.fpfh, cores)
            vals.append(v); t_all += dt
        val = float(np.mean(vals))
        sample = f"{n} pairs of the workload per step (seed {args.seed}), {cores} worker processes"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * t_all / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
                          "data": "synthetic", "config": config_dict(args, {"pairs_per_step": n}),
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream(dev)
    eng = g.Engine(local, stream.cuda_stream)          # kernels run on torch's current stream: torch events see them
    eng.set_options(args.exact, args.spec, -1)
    eng.set_batch_options(args.groups, args.slots)
    kw = dict(cfpfh=1, regularizationFPFH=0.000005) if args.fpfh else {}
    params = g.shipped_config(**kw)
    pairs = synth.bo1_pairs(args.pairs, seed=args.seed + 7919 * rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def evals_of(res):
        return float(sum(r["counters"][2] * p["nd"] for r, p in zip(res, pairs)))

    # ---- resident leg: inputs uploaded before the timed region ----
    eng.batch_upload(params, pairs)
    for _ in range(args.warmup):
        flush.zero_()
        res = eng.batch_run()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms, launches, total_evals = 0.0, 0, 0.0
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        res = eng.batch_run()
        ev[k][1].record(stream)
        tm = eng.timings()
        kern_ms += tm["ms"][2]; launches += sum(tm["launches"]); total_evals += evals_of(res)
        last_tm, last_stats = tm, eng.stats()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # ---- e2e leg: host buffers in, results out, through the public C-ABI call ----
    descs = eng.make_descs(pairs)   # goicp_pair_desc[]: pointers to the host (numpy) buffers, built once like any C caller would
    for _ in range(min(args.warmup, 1)):
        eng.register_batch_descs(params, descs)
    barrier()
    e2e_ms, e2e_evals = 0.0, 0.0
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        r2 = eng.register_batch_descs(params, descs)   # host buffers -> H2D -> DT build -> Register -> results in host memory
        torch.cuda.synchronize(dev)
        e2e_ms += 1e3 * (time.perf_counter() - t1); e2e_evals += float(sum(r.counters[2] * p["nd"] for r, p in zip(r2, pairs)))
    barrier()
    sampler.stop_flag = True
    h2d = sum(p[k].nbytes for p in pairs for k in ("model_xyz", "data_xyz", "model_c", "data_c") ) + (sum(p["model_fpfh"].nbytes + p["data_fpfh"].nbytes for p in pairs) if args.fpfh else 0)
    d2h = len(pairs) * 208

    stats = torch.tensor([dev_ms, e2e_ms, total_evals, e2e_evals, kern_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms, kern_ms = mx[0].item(), mx[1].item(), mx[4].item()
        total_evals, e2e_evals, launches = sm[2].item(), sm[3].item(), int(sm[5].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = total_evals / (dev_ms * 1e-3)
    peaks, peak_src = {"hbm_gbs": 6650.0}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))); peak_src = "measured"
    except Exception:
        pass
    kern_evals_per_gpu = total_evals / world
    # 4 B per cube.point eval (one DT voxel) + with regularization>0 27/8 corner evals x (4 B index map + 2 B
    # property mask) = 20.25 B (+ 27/8 x 4 B of the per-cell c-FPFH table with the c-FPFH term)
    bytes_per_eval = 4.0 + (20.25 if params.regularization > 0 else 0.0) + (13.5 if args.fpfh else 0.0)
    achieved = kern_evals_per_gpu * bytes_per_eval / (kern_ms * 1e-3) / 1e9
    agg = total_evals / world * bytes_per_eval / (dev_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "inner_bnb_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": None, "peak_source": peak_src, "achieved_aggregate": agg, "frac_aggregate": agg / peaks["hbm_gbs"],
            "algorithmic_bytes_per_eval": bytes_per_eval, "achieved_dt_gather_only": kern_evals_per_gpu * 4 / (kern_ms * 1e-3) / 1e9,
            "note": "algorithmic bytes per eval per SURVEY 8(d): 4 B DT voxel + 20.25 B corner-term gathers when regularization>0. `achieved` = evals x bytes / CUDA-event time of the resident inner_bnb kernel (one launch per step serves every InnerBnB and "
                    "ICP request of the batch), `achieved_aggregate` = evals x bytes / step time per GPU, `achieved_dt_gather_only` counts the 4 B DT voxel alone. Only evals of calls the reference order consumes are counted (speculative "
                    "calls and corner evals are not). The 20^3 volumes are staged per call into shared memory by TMA (cp.async.bulk), so the gathers are LDS and DRAM sees only the staging traffic: the binding limit is SM issue + the serial "
                    "phases of each queue pop, not HBM -- see DESIGN.md section 4. `traffic` = DRAM bytes of one classic-mode launch of the same kernel (ncu cannot replay the resident kernel; profiles/README.md)"}
    import glob
    profs = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_inner_bnb_traffic.json")))
    prof = profs[-1] if profs else ""
    if prof and os.path.exists(prof):
        try:
            roof["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": config_dict(args), "pairs_per_s": args.pairs * world * args.steps / (dev_ms * 1e-3), "ms_per_pair": dev_ms / args.steps / args.pairs,
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "e2e": {"value": e2e_evals / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "pairs_per_s": args.pairs * world * args.steps / (e2e_ms * 1e-3)},
            "gpu_launches": int(launches), "roofline": roof, "clocks": sampler.summary(),
            "rank0_last_step": {"gpu_ms_sum_over_streams": {"dt": last_tm["ms"][0], "initialize": last_tm["ms"][1], "inner_bnb": last_tm["ms"][2], "icp": last_tm["ms"][3]},
                                "launches": last_tm["launches"], **last_stats}}
    if not args.no_cpu_baseline and world == 1:
        n = args.cpu_pairs or max(8, 2 * cores)
        v, dt, kind, errs = cpu_arm(pairs[:n], args.fpfh, cores)
        same = all(np.float32(e) == np.float32(r["optError"]) for e, r in zip(errs, res[:n]))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "seconds": dt,
                                "sample": f"the first {n} pairs of the step's workload, {cores} worker processes (one reference process per pair)",
                                "opt_error_identical_to_gpu": bool(same)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
