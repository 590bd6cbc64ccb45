"""Print selected fields of the bench JSON line in a file that may also hold other stdout lines (NCCL banner)."""
import json
import sys

line = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if not line:
    sys.exit(1)
d = json.loads(line[-1])
print(d.get("value"), d.get("unit"), d.get("ms_per_step"), d.get("scaling"), d.get("config", {}).get("global_batch"),
      d.get("config", {}).get("workload", "")[:60], d.get("grad_allreduce"), d.get("peak_mem_gb"))
