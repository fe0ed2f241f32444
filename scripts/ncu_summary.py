"""Summarises an ncu launch list and full ncu reports from gpurun_out/ into profiles/ (tracked).
    python scripts/ncu_summary.py <tag> [<report base>:<label> ...]
e.g. python scripts/ncu_summary.py r02F search:search bunny300_inner:bunny300_inner_bnb
reads  gpurun_out/<tag>_launches.csv (if present) and gpurun_out/<tag>_<report base>.ncu-rep
writes profiles/<tag>_launches_summary.csv, profiles/<tag>_<label>_{raw.csv,traffic.json,sass_mix.csv,stalls.csv}"""
import collections, csv, json, os, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02F"
reports = [a.split(":") for a in sys.argv[2:]]
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

lf = f"gpurun_out/{tag}_launches.csv"
if os.path.exists(lf):
    rows = list(csv.reader(open(lf)))
    hdr = None; agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r: hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                name = d["Kernel Name"].split("(")[0]; v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
                v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
                agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, cold-cache and serialised -- compare SHARES", "# command: see profiles/README.md", "kernel,launches,total_us,share,avg_us"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): out.append(f"\"{k}\",{v[0]},{v[1]:.1f},{v[1]/tot:.4f},{v[1]/v[0]:.1f}")
    open(f"profiles/{tag}_launches_summary.csv", "w").write("\n".join(out) + "\n")
    print("\n".join(out))

want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_lsu.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for base, label in reports:
    rep = f"gpurun_out/{tag}_{base}.ncu-rep"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines())); h = rr[0]
    idx = [i for i, x in enumerate(h) if x in want]
    lines = [",".join(h[i] for i in idx), ",".join(rr[1][i] for i in idx)] + [",".join('"%s"' % r[i] if "," in r[i] else r[i] for i in idx) for r in rr[2:]]
    open(f"profiles/{tag}_{label}_raw.csv", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    num = lambda s: float(s.replace(",", ""))
    dr = [num(r[h.index("dram__bytes_read.sum")]) for r in rr[2:]]; du = rr[1][h.index("dram__bytes_read.sum")]
    dw = [num(r[h.index("dram__bytes_write.sum")]) for r in rr[2:]]; dwu = rr[1][h.index("dram__bytes_write.sum")]
    traffic = sum(a * mul[du] + b * mul[dwu] for a, b in zip(dr, dw)) / len(dr)
    json.dump({"kernel": label, "dram_bytes_per_launch": traffic, "dram_bytes_read": sum(dr) * mul[du] / len(dr), "dram_bytes_write": sum(dw) * mul[dwu] / len(dw), "launches_profiled": len(dr),
               "source": f"profiles/{tag}_{label}_raw.csv (ncu --set full --clock-control none)"}, open(f"profiles/{tag}_{label}_traffic.json", "w"))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines())); hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]; hd = rows[hi]; ix = {x: i for i, x in enumerate(hd)}
    ops = collections.Counter(); samp = collections.Counter(); tot = tots = 0
    stall_cols = [c for c in hd if c.startswith("stall_") and "Not Issued" not in c]; stalls = collections.Counter()
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name": break
        if len(r) < len(hd) or not r[0].startswith("0x"): continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip()); op = m.group(2).split(".")[0] if m else "?"
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]]); ops[op] += n; samp[op] += s; tot += n; tots += s
        for c in stall_cols: stalls[c] += int(r[ix[c]])
    lines = [f"# SASS opcode mix of {label} (first profiled launch): warp instructions executed, share, share of warp-state samples", f"total,{tot},1.0,{tots}"]
    lines += [f"{op},{n},{n/tot:.4f},{samp[op]/max(tots,1):.4f}" for op, n in ops.most_common(30)]
    open(f"profiles/{tag}_{label}_sass_mix.csv", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:16]))
    st = sum(stalls.values())
    lines = [f"# warp-state samples of {label} by reason (source page, all samples), share"] + [f"{c},{n},{n/max(st,1):.4f}" for c, n in stalls.most_common()]
    open(f"profiles/{tag}_{label}_stalls.csv", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))
