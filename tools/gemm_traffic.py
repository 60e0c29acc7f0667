"""DRAM traffic of the GEMM launches in an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` launch list of one pre-training step -> JSON (bench.py's roofline.traffic reads it from profiles/)."""
import csv
import json
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
per = defaultdict(dict)
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "")
    if r["Metric Name"] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per[r["ID"]]["name"] = r["Kernel Name"]
    per[r["ID"]][r["Metric Name"]] = v
g = [k for k in per.values() if "gemm_tc_kernel" in k["name"]]
byt = lambda k: k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)  # noqa: E731
res = {
    "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one "
              "pre-training step (tools/profile_step.py)",
    "gemm_launches": len(g),
    "gemm_time_us": sum(k["gpu__time_duration.sum"] for k in g),
    "gemm_dram_bytes": sum(byt(k) for k in g),
    "gemm_dram_bytes_per_launch": sum(byt(k) for k in g) / max(1, len(g)),
    "all_kernels_time_us": sum(k["gpu__time_duration.sum"] for k in per.values()),
    "all_kernels_dram_bytes": sum(byt(k) for k in per.values()),
}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
