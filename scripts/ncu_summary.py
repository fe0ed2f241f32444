"""Summarises gpurun_out/r01_launches.csv (ncu launch list) and a full ncu report into profiles/ (tracked)."""
import collections, csv, json, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = list(csv.reader(open(f"gpurun_out/{tag}_launches.csv")))
hdr = None; agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            name = d["Kernel Name"].split("(")[0]; v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
            v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
            agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out = [f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, cold-cache and serialised -- compare SHARES", f"# command: see profiles/README.md", "kernel,launches,total_us,share,avg_us"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): out.append(f"{k},{v[0]},{v[1]:.1f},{v[1]/tot:.4f},{v[1]/v[0]:.1f}")
open(f"profiles/{tag}_launches_summary.csv", "w").write("\n".join(out) + "\n")
print("\n".join(out))
raw = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_inner_bnb.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum"]
idx = [i for i, x in enumerate(h) if x in want]
lines = [",".join(h[i] for i in idx), ",".join(rr[1][i] for i in idx)] + [",".join(r[i] for i in idx) for r in rr[2:]]
open(f"profiles/{tag}_inner_bnb_raw.csv", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
dr = [float(r[h.index("dram__bytes_read.sum")]) for r in rr[2:]]; du = rr[1][h.index("dram__bytes_read.sum")]
dw = [float(r[h.index("dram__bytes_write.sum")]) for r in rr[2:]]; dwu = rr[1][h.index("dram__bytes_write.sum")]
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = sum(a * mul[du] + b * mul[dwu] for a, b in zip(dr, dw)) / len(dr)
json.dump({"kernel": "inner_bnb_kernel", "dram_bytes_per_launch": traffic, "launches_profiled": len(dr), "source": f"profiles/{tag}_inner_bnb_raw.csv (ncu --set full)"}, open(f"profiles/{tag}_inner_bnb_traffic.json", "w"))
src = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_inner_bnb.ncu-rep", "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]; hd = rows[hi]; ix = {x: i for i, x in enumerate(hd)}
ops = collections.Counter(); samp = collections.Counter(); tot = tots = 0
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name": break
    if len(r) < len(hd) or not r[0].startswith("0x"): continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip()); op = m.group(2).split(".")[0] if m else "?"
    n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]]); ops[op] += n; samp[op] += s; tot += n; tots += s
lines = [f"# SASS opcode mix of inner_bnb_kernel<EXACT> (first profiled launch): warp instructions executed, share, share of stall samples", f"total,{tot},1.0,{tots}"]
lines += [f"{op},{n},{n/tot:.4f},{samp[op]/max(tots,1):.4f}" for op, n in ops.most_common(30)]
open(f"profiles/{tag}_inner_bnb_sass_mix.csv", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:20]))
