"""Builds libccj_b200.so (CUDA, sm_100a), the CCJ command line and the oracle binaries, all in-tree.

    python -m ccj_b200.build            # library + CLI (+ oracle when possible)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "ccj_b200" / "csrc"
LIB = ROOT / "ccj_b200" / "libccj_b200.so"
CLI = ROOT / "ccj_b200" / "bin" / "CCJ"

# the C++ class shells (W_final / pseudo_loop / s_energy_matrix on top of the C ABI); host code, g++
SHELL_SOURCES = [CSRC / "W_final.cc", CSRC / "ccj_classes.cc", CSRC / "ccj_shell.cc"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr", "-diag-suppress", "20012",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required, there is no CPU fallback")


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, **kw):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True, **kw)


def embedded_par_defines():
    """-D flags that tell embedded_params.cpp where the reference's parameter data files are (.incbin)."""
    par = ROOT / "params"
    return [f'-DCCJ_PAR_TURNER04="{par / "rna_Turner04.par"}"', f'-DCCJ_PAR_DNA_MATHEWS04="{par / "dna_Matthews04.par"}"']


def build_library(force: bool = False, verbose_ptxas: bool = False) -> Path:
    sources = [CSRC / "ccj_abi.cu", CSRC / "ccj_kernels.cu", CSRC / "ccj_fill4.cu", CSRC / "ccj_peak.cu", CSRC / "ccj_shard.cu",
               CSRC / "energy_model.cpp", CSRC / "embedded_params.cpp"]
    sources = [s for s in sources if s.exists()]
    deps = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.hpp")) + sources + [
        ROOT / "include" / "ccj_b200.h", ROOT / "params" / "rna_Turner04.par", ROOT / "params" / "dna_Matthews04.par"]
    if not force and not _newer(LIB, deps):
        return LIB
    flags = list(NVCC_FLAGS) + embedded_par_defines()
    if verbose_ptxas:
        flags += ["-Xptxas", "-v"]
    _run([_nvcc(), *flags, "-shared", "-o", LIB, *sources, "-I", ROOT / "include", "-lcudart", "-ldl"])
    return LIB


def build_cli(force: bool = False) -> Path:
    sources = [CSRC / "CCJ.cc", CSRC / "cmdline.cc", *SHELL_SOURCES]
    if not all(s.exists() for s in sources):
        return CLI
    deps = sources + list(CSRC.glob("*.hh")) + list(CSRC.glob("*.hpp")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [LIB]
    if not force and not _newer(CLI, deps):
        return CLI
    CLI.parent.mkdir(parents=True, exist_ok=True)
    _run(["g++", "-std=c++17", "-O2", "-o", CLI, *sources, "-I", ROOT / "include", "-I", CSRC,
          "-L", LIB.parent, "-lccj_b200", f"-Wl,-rpath,$ORIGIN/..", ])
    return CLI


def build_oracle() -> None:
    """Builds oracle/_ref (the unmodified reference, only where /root/reference exists) and the
    CPU restatement. Building the checker is not using it: nothing in the product links it."""
    odir = ROOT / "oracle"
    targets = []
    if (odir / "ccj_oracle.cc").exists():
        targets.append("_ref/ccj_oracle")
    if Path(os.environ.get("CCJ_REFERENCE", "/root/reference")).exists():
        targets.append("ref")
        if LIB.exists():
            targets.append("refmain")   # the reference's own main() against the B200 shells (test tool)
    if targets:
        _run(["make", "-C", odir, "-j8", *targets], stdout=subprocess.DEVNULL)


def build_all(force: bool = False) -> None:
    build_library(force)
    build_cli(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
