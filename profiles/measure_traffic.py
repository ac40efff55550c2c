"""Sums the DRAM traffic ncu measured for every gap-table launch of traffic_workload.py and sets it against the
algorithmic bytes of the same fold:  python profiles/measure_traffic.py gpurun_out/traffic.csv [count] [out.json]
Writes profiles/r2_traffic.json (read by bench.py for roofline.traffic) and a per-kernel summary."""
import csv
import json
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import ccj_b200  # noqa: E402

path = sys.argv[1]
count = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1}
per = defaultdict(lambda: defaultdict(float))
launches = defaultdict(int)
for r in rows[1:]:
    name = r[ik].split("(")[0].replace("void ", "")
    v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1)
    per[name][r[im]] += v
    if r[im] == "gpu__time_duration.sum":
        launches[name] += 1
GAP = ("k_roles", "k_winLR", "k_winM", "k_final")
gap = {k: v for k, v in per.items() if k.startswith(GAP)}
dram = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in gap.values())
alg = sum(ccj_b200.count_terms(s)["bytes_4d"] for s in bench.workload(0, count))
out = {"dram_bytes": dram, "algorithmic_bytes": alg, "ratio": dram / alg,
       "source": f"ncu dram__bytes_read+write over all {sum(launches[k] for k in gap)} gap-table launches of one fill of "
                 f"{count} x 150-nt workload sequence(s) (profiles/traffic_workload.py)",
       "per_kernel": {k: {"launches": launches[k], "dram_read_GB": v["dram__bytes_read.sum"] / 1e9,
                          "dram_write_GB": v["dram__bytes_write.sum"] / 1e9,
                          "time_ms_under_ncu": v["gpu__time_duration.sum"] * 1e3} for k, v in per.items()}}
out_name = sys.argv[3] if len(sys.argv) > 3 else "r2_traffic.json"
(ROOT / "profiles" / out_name).write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
