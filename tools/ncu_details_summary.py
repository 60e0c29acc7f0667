"""`ncu --page details --csv` -> a markdown table of the headline metrics per kernel (profiles/*_ncu_*.md)."""
import csv
import sys
from collections import OrderedDict

WANT = [("GPU Speed Of Light Throughput", "Duration"), ("GPU Speed Of Light Throughput", "DRAM Throughput"),
        ("GPU Speed Of Light Throughput", "Compute (SM) Throughput"), ("GPU Speed Of Light Throughput", "L2 Cache Throughput"),
        ("Memory Workload Analysis", "Memory Throughput"), ("Memory Workload Analysis", "L2 Hit Rate"),
        ("Compute Workload Analysis", "Issue Slots Busy"), ("Compute Workload Analysis", "Executed Ipc Active"),
        ("Occupancy", "Achieved Occupancy"), ("Launch Statistics", "Registers Per Thread"),
        ("Launch Statistics", "Grid Size"), ("Launch Statistics", "Block Size"),
        ("Launch Statistics", "Dynamic Shared Memory Per Block"),
        ("Scheduler Statistics", "Eligible Warps Per Scheduler"), ("Warp State Statistics", "Warp Cycles Per Issued Instruction")]
for path in sys.argv[1:]:
    rows = list(csv.DictReader(open(path)))
    byk = OrderedDict()
    for r in rows:
        byk.setdefault((r["ID"], r["Kernel Name"]), {})[(r["Section Name"], r["Metric Name"])] = (r["Metric Value"], r["Metric Unit"])
    print("### %s\n" % path)
    print("| kernel | " + " | ".join(m for _, m in WANT) + " |")
    print("|---|" + "---:|" * len(WANT))
    for (_, name), v in byk.items():
        cells = []
        for w in WANT:
            val = v.get(w)
            cells.append("%s %s" % val if val else "-")
        print("| `%s` | %s |" % (name[:70], " | ".join(cells)))
    print()
