"""Shared fixtures.  Markers: `gpu` = needs a B200 (the parity tests proper, through the C ABI)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"
REF_BIN = ROOT / "oracle" / "_ref" / "CCJ"
REF_DUMP = ROOT / "oracle" / "_ref" / "ccj_ref_dump"
EMU_BIN = ROOT / "build" / "ccj_emu"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def library():
    """The in-tree CUDA library (built on demand; nvcc cross-compiles without a GPU)."""
    from ccj_b200 import build
    import ccj_b200
    build.build_library()
    return ccj_b200.load_library()


@pytest.fixture(scope="session")
def emu_bin():
    """Host build of the product's cell functions (tests/emu/ccj_emu.cpp), g++ only."""
    srcs = [ROOT / "tests" / "emu" / "ccj_emu.cpp", ROOT / "ccj_b200" / "csrc" / "energy_model.cpp",
            ROOT / "ccj_b200" / "csrc" / "embedded_params.cpp"]
    deps = srcs + list((ROOT / "ccj_b200" / "csrc").glob("*.cuh")) + list((ROOT / "ccj_b200" / "csrc").glob("*.h*"))
    if not EMU_BIN.exists() or any(d.stat().st_mtime > EMU_BIN.stat().st_mtime for d in deps):
        EMU_BIN.parent.mkdir(exist_ok=True)
        from ccj_b200 import build
        subprocess.run(["g++", "-std=c++17", "-O2", *build.embedded_par_defines(), "-o", str(EMU_BIN)] +
                       [str(s) for s in srcs], check=True)
    return EMU_BIN


@pytest.fixture(scope="session")
def golden_folds():
    return json.loads((GOLDEN / "folds.json").read_text())


@pytest.fixture(scope="session")
def golden_hashes():
    return json.loads((GOLDEN / "table_hashes.json").read_text())


@pytest.fixture(scope="session")
def golden_long():
    p = GOLDEN / "folds_long.json"
    return json.loads(p.read_text()) if p.exists() else []


def par_path(name):
    return str(ROOT / "params" / name)


@pytest.fixture(scope="session")
def ctx_factory(library):
    import ccj_b200
    made = {}

    def get(par="rna_Turner04.par", dangles=2, no_gu=False):
        key = (par, dangles, no_gu)
        if key not in made:
            made[key] = ccj_b200.Context(0, par_path(par), dangles, no_gu)
        return made[key]

    yield get
    for c in made.values():
        c.close()
