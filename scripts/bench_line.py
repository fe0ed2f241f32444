"""one-line summary of a bench.py log: python scripts/bench_line.py <log file> [tag]   (reads the file, never stdin)"""
import json, sys
path = sys.argv[1]; tag = sys.argv[2] if len(sys.argv) > 2 else path
try:
    d = [json.loads(l) for l in open(path) if l.startswith("{")][-1]
    print(tag, "evals/s %.3g pairs/s %.1f ms/step %.1f e2e pairs/s %.1f" % (d["value"], d["pairs_per_s"], d["ms_per_step"], d["e2e"]["pairs_per_s"]), d.get("rank0_last_step"))
except Exception as e:
    print(tag, "FAILED", e)
