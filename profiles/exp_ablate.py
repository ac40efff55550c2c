"""Runs profiles/exp_variant.py once per library variant in ccj_b200/variants (built by profiles/build_variants.sh).
usage: python profiles/exp_ablate.py [--only-profiled] name ...   -> one JSON line per variant"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
args = sys.argv[1:]
only = "--only-profiled" in args
names = [a for a in args if not a.startswith("--")]
for name in names:
    env = dict(os.environ, CCJ_B200_LIB=str(ROOT / "ccj_b200" / "variants" / f"libccj_{name}.so"))
    if only:
        env["CCJ_EXP_PROFILED_ONLY"] = "1"
    r = subprocess.run([sys.executable, str(ROOT / "profiles" / "exp_variant.py")], env=env, capture_output=True, text=True, timeout=300)
    print(name, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else f"rc={r.returncode} {r.stderr[-300:]}", flush=True)
