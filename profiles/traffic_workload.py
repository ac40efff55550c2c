"""One fill of the first sequences of the benchmark workload (config 4), for ncu:
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
       -k regex:'k_roles|k_win|k_final' --csv --log-file gpurun_out/traffic.csv python profiles/traffic_workload.py 1
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import ccj_b200  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 1
seqs = bench.workload(0, count)
with ccj_b200.Context(0, ccj_b200.default_par_file(bench.PAR), 2) as ctx:
    ctx.prepare(seqs)
    print("fill ms", ctx.fill())
