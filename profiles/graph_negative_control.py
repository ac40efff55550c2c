"""Negative control of the launch-graph check (tests/test_emu_tuned.py::test_launch_graph_of_enqueue_fill_...): every
cudaStreamWaitEvent of enqueue_fill() is dropped in turn and the emulated kernels run in the six stream-priority orders;
a dropped wait is "detected" when some order gives wrong tables.  CPU only: python profiles/graph_negative_control.py
(needs build/ccj_emu_graph and build/ccj_emu from a run of the CPU suite)."""
import itertools, os, subprocess, sys, time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
seq = "UUGUCAGAACGCUG"
par = str(ROOT / "params" / "rna_Turner04.par")
ref = subprocess.run([str(ROOT / "build" / "ccj_emu"), "hash", par, "2", seq], capture_output=True, text=True, check=True).stdout
orders = ["prio:" + "".join(p) for p in itertools.permutations("123")]


def run(k, order, cores):
    env = dict(os.environ, CCJ_EMU_POISON="80")
    if k >= 0:
        env["CCJ_EMU_DROP_WAIT"] = str(k)
    p = subprocess.run([str(ROOT / "build" / "ccj_emu_graph"), "hash", par, "2", seq, "0", "-1", "tuned", order], capture_output=True,
                       text=True, env=env, preexec_fn=lambda: os.sched_setaffinity(0, cores))
    return k, order, p.stdout == ref, (p.stderr.strip().splitlines() or [""])[0]


jobs = [(k, o) for k in range(-1, 41) for o in orders]
res = {}
t0 = time.time()
with ThreadPoolExecutor(4) as ex:
    for k, o, ok, msg in [f.result() for f in [ex.submit(run, k, o, {2 * (i % 4), 2 * (i % 4) + 1}) for i, (k, o) in enumerate(jobs)]]:
        res.setdefault(k, [msg]).append((o, ok))
det = 0
for k in sorted(res):
    bad = [o[5:] for o, ok in res[k][1:] if not ok]
    det += bool(bad) and k >= 0
    print(k, res[k][0][:110], "| wrong tables in orders " + ",".join(bad) if bad else "| every order right")
print(f"{det} of 41 dropped waits give wrong tables in some order ({time.time() - t0:.0f} s); streams: 1 main, 2 windows, 3 P + 2D")
