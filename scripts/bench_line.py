import json, sys
tag = sys.argv[1]
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print(tag, "evals/s %.3g pairs/s %.1f ms/step %.1f e2e pairs/s %.1f" % (d["value"], d["pairs_per_s"], d["ms_per_step"], d["e2e"]["pairs_per_s"]), d["rank0_last_step"])
except Exception as e:
    print(tag, "FAILED", e)
