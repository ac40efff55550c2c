"""Python front end of the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / reference arm, never by ccj_b200/).

  ccj_oracle  the CPU restatement (oracle/ccj_oracle.cc): table hashes and W[n] for one sequence
  _ref/CCJ    the unmodified reference, compiled by oracle/Makefile where /root/reference exists
"""
from __future__ import annotations

import gzip
import subprocess
import tempfile
from pathlib import Path
from typing import Dict, List, Optional

ODIR = Path(__file__).resolve().parent
ROOT = ODIR.parent
ORACLE_BIN = ODIR / "_ref" / "ccj_oracle"
REF_BIN = ODIR / "_ref" / "CCJ"
REF_DUMP = ODIR / "_ref" / "ccj_ref_dump"
_params_cache: Dict[str, str] = {}


def build_oracle() -> Path:
    src = ODIR / "ccj_oracle.cc"
    if not ORACLE_BIN.exists() or ORACLE_BIN.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ODIR), "_ref/ccj_oracle"], check=True, stdout=subprocess.DEVNULL)
    return ORACLE_BIN


def params_dump(par_name: str) -> str:
    """Unpacked copy of the scaled-parameter dump of `par_name` that the compiled reference wrote
    (tests/golden/params_<name>.txt.gz, see tests/golden/make_golden.py)."""
    stem = Path(par_name).stem
    if stem not in _params_cache:
        gz = ROOT / "tests" / "golden" / f"params_{stem}.txt.gz"
        if not gz.exists():
            raise FileNotFoundError(f"no golden parameter dump for {par_name}")
        tmp = tempfile.NamedTemporaryFile(prefix=f"ccj_params_{stem}_", suffix=".txt", delete=False)
        tmp.write(gzip.decompress(gz.read_bytes()))
        tmp.close()
        _params_cache[stem] = tmp.name
    return _params_cache[stem]


def _run(mode: str, seq: str, par_name: str, dangles: int, no_gu: bool) -> str:
    exe = build_oracle()
    p = subprocess.run([str(exe), mode, params_dump(par_name), str(dangles), "1" if no_gu else "0", seq],
                       capture_output=True, text=True, check=True)
    return p.stdout


def oracle_energy_dcal(seq: str, par_name: str = "rna_Turner04.par", dangles: int = 2, no_gu: bool = False) -> int:
    """W[n] of the restated fill, in dcal/mol (the reference prints W[n]/100)."""
    return int(_run("energy", seq, par_name, dangles, no_gu).strip())


def oracle_hashes(seq: str, par_name: str = "rna_Turner04.par", dangles: int = 2, no_gu: bool = False):
    """({table: [finite, min|sum, fnv]}, W[n]) in the format of tests/golden/table_hashes.json."""
    tables, w = {}, None
    for line in _run("hash", seq, par_name, dangles, no_gu).splitlines():
        f = line.split()
        if f[0] == "n":
            continue
        if f[0] == "W":
            w = int(f[1])
        else:
            tables[f[0]] = [int(f[1]), int(f[2]), f[3]]
    return tables, w


def reference_fold(seq: str, par_file: str, extra: Optional[List[str]] = None):
    """(returncode, stdout, stderr) of the unmodified reference binary, or None where it did not travel."""
    if not REF_BIN.exists():
        return None
    p = subprocess.run([str(REF_BIN), "-P", par_file, *(extra or []), seq], capture_output=True, text=True)
    return p.returncode, p.stdout, p.stderr


def oracle_window_terms(seq: str, par_name: str = "rna_Turner04.par", dangles: int = 2, no_gu: bool = False) -> dict:
    """Interior-window candidates the restated fill evaluates: {"iloop": PL+PR+PM, "PL", "PR", "PM", "PO_dead"}."""
    f = _run("count", seq, par_name, dangles, no_gu).split()
    return {f[x]: int(f[x + 1]) for x in range(0, len(f), 2)}
