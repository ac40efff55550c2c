"""The C++ class surface: W_final / pseudo_loop / s_energy_matrix with the reference's constructors,
public members and methods (src/W_final.hh:18-71, src/pseudo_loop.hh:13-56, src/s_energy_matrix.hh:16-68).

  * CPU: struct layouts of ccj_compat.hh == the reference's headers (offset for offset), and a translation unit
    written against the reference's public API compiles and links against the shells.
  * GPU: tests/shell/probe_body.inc -- ONE source compiled against both header sets -- prints every getter, the
    loop-energy helpers and a node-by-node pseudo_loop::backtrack; the text must equal what the unmodified reference
    printed (tests/golden/probe_*.txt.gz, and live oracle/_ref/ccj_ref_dump where it travelled)."""
import gzip
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))
CSRC = ROOT / "ccj_b200" / "csrc"
BUILD = ROOT / "build"


def compile_against_shells(src, out, link=True):
    from ccj_b200 import build
    BUILD.mkdir(exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", str(CSRC), "-I", str(ROOT / "tests" / "shell")]
    if link:
        cmd += ["-o", str(out), str(src), *[str(s) for s in build.SHELL_SOURCES], "-L", str(ROOT / "ccj_b200"), "-lccj_b200",
                f"-Wl,-rpath,{ROOT / 'ccj_b200'}"]
    else:
        cmd += ["-o", str(out), str(src)]
    subprocess.run(cmd, check=True)
    return out


def test_struct_layouts_equal_the_reference_headers():
    exe = compile_against_shells(ROOT / "tests" / "shell" / "layout_probe.cc", BUILD / "layout_probe", link=False)
    mine = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    assert mine == (ROOT / "tests" / "golden" / "vrna_layout.txt").read_text()
    live = ROOT / "oracle" / "_ref" / "layout_probe_ref"
    if live.exists():
        assert mine == subprocess.run([str(live)], capture_output=True, text=True, check=True).stdout


@pytest.fixture(scope="module")
def shell_probe(library):
    return compile_against_shells(ROOT / "tests" / "shell" / "shell_probe.cc", BUILD / "shell_probe")


def test_reference_style_code_compiles_and_links(shell_probe):
    assert shell_probe.exists()


@pytest.mark.gpu
def test_class_surface_matches_the_reference(shell_probe):
    from make_golden import PROBES
    dump = ROOT / "oracle" / "_ref" / "ccj_ref_dump"
    for name, par, d, seq in PROBES:
        p = subprocess.run([str(shell_probe), str(ROOT / "params" / par), str(d), seq], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        with gzip.open(ROOT / "tests" / "golden" / f"probe_{name}.txt.gz", "rt") as f:
            want = f.read()
        got = p.stdout
        if got != want:
            for a, b in zip(got.splitlines(), want.splitlines()):
                assert a == b, name
        assert got == want, name
        if dump.exists() and name.startswith("r38"):
            live = subprocess.run([str(dump), "probe", str(ROOT / "params" / par), str(d), seq], capture_output=True, text=True)
            assert live.stdout == got


@pytest.mark.gpu
def test_table_getters_of_the_c_abi(ctx_factory):
    """ccj_table4_get / ccj_table2_get: Matrix4D::get semantics (INF for an invalid index, src/matrices.hh:177-182)."""
    import ctypes as C
    import ccj_b200
    ctx = ctx_factory()
    seq = "GGGAAACGCUCUAGCGUUUCCCAAAGAGCAAAUCGAUCA"
    ctx.prepare([seq])
    ctx.fill()
    n = len(seq)
    lib = ctx._lib
    v = C.c_int32()
    for name in ("PK", "PL", "PfromO", "POmloop10"):
        t = ccj_b200.TABLE4.index(name)
        full = ctx.table4(0, name)
        x = 0
        for i in range(1, n + 1):
            for j in range(i, n + 1):
                for k in range(j + 2, n + 1):
                    for l in range(k, n + 1):
                        if x % 53 == 0:
                            assert lib.ccj_table4_get(ctx._h, 0, t, i, j, k, l, C.byref(v)) == 0
                            assert v.value == int(full[x])
                        x += 1
        for bad in ((3, 2, 6, 7), (1, 4, 5, 9), (0, 1, 4, 5), (1, 2, 5, n + 1), (2, 2, 3, 3)):
            assert lib.ccj_table4_get(ctx._h, 0, t, *bad, C.byref(v)) == 0
            assert v.value == 10000000
    for name in ("V", "WM", "P", "WBP", "WPP"):
        t = ccj_b200.TABLE2.index(name)
        full = ctx.table2(0, name)
        x = 0
        for i in range(1, n + 1):
            for j in range(i, n + 1):
                assert lib.ccj_table2_get(ctx._h, 0, t, i, j, C.byref(v)) == 0
                assert v.value == int(full[x])
                x += 1
        assert lib.ccj_table2_get(ctx._h, 0, t, 3, 2, C.byref(v)) != 0      # i > j is the caller's case (reference: return_val)
        assert lib.ccj_table2_get(ctx._h, 0, t, 1, n + 1, C.byref(v)) != 0
