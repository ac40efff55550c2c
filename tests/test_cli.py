"""The CCJ command line (ccj_b200/bin/CCJ) is a drop-in for the reference binary: option handling is
checked on the CPU (against oracle/_ref/CCJ when it is present, else against the recorded behaviour),
folds on the GPU against the golden vectors."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CLI = ROOT / "ccj_b200" / "bin" / "CCJ"
REF = ROOT / "oracle" / "_ref" / "CCJ"

HELP_HEAD = "Usage: CCJ [options] [sequence]\nPseudoknotted minimum free energy folding of RNAs\n"


@pytest.fixture(scope="module")
def cli(library):
    from ccj_b200 import build
    build.build_cli()
    assert CLI.exists()
    return CLI


def run(exe, args, stdin=None, cwd=ROOT):
    p = subprocess.run([str(exe)] + list(args), input=stdin, capture_output=True, text=True, cwd=str(cwd))
    # the program name in getopt messages is argv[0]: normalise it
    return p.returncode, p.stdout, p.stderr.replace(str(exe), "CCJ")


OPTION_CASES = [["--help"], ["-h"], ["-V"], ["--version"], ["--foo"], ["-d"], ["-d", "x", "ACGU"], ["-d1", "-d2", "ACGU"],
                ["-P", "/nonexistent", "ACGU"], ["-i", "somefile"], ["ACGN"], ["--noGU", "--noGU", "ACGU"]]


@pytest.mark.parametrize("args", OPTION_CASES)
def test_option_handling_matches_reference(cli, args):
    mine = run(cli, args)
    if REF.exists():
        assert mine == run(REF, args), args
    if args[0] in ("--help", "-h"):
        assert mine[0] == 0 and mine[1].startswith(HELP_HEAD) and "--noGU" in mine[1]
    if args[0] in ("-V", "--version"):
        assert mine == (0, "CCJ 1.0\n", "")
    if args == ["--foo"]:
        assert mine[0] == 1 and "unrecognized option '--foo'" in mine[2]
    if args == ["-P", "/nonexistent", "ACGU"]:
        assert mine == (1, "", "Not a valid parameter file!\n")
    if args == ["-i", "somefile"]:
        assert mine == (1, "sequence is missing\n", "")
    if args == ["ACGN"]:
        assert mine == (1, "Sequence contains character N that is not G,C,A,U, or T.\n", "")


def test_empty_stdin(cli):
    assert run(cli, [], stdin="\n") == (1, "sequence is missing\n", "")


@pytest.mark.gpu
def test_cli_folds_match_golden(cli, golden_folds):
    recs = [r for r in golden_folds if len(r["seq"]) <= 60][:40] + [r for r in golden_folds if r["rc"] != 0][:4]
    for r in recs:
        args = ["-P", str(ROOT / "params" / r["par"]), "-d", str(r["dangles"])] + r["extra"] + [r["seq"]]
        assert run(cli, args) == (r["rc"], r["stdout"], r["stderr"]), r["seq"]


@pytest.mark.gpu
def test_cli_input_conventions(cli):
    """stdin input, lower case, T->U, default parameter file relative to the cwd, extra positionals ignored."""
    want = "GCAACGAUGACAUACAUCGCUAGUCGACGC\n....(((((.....)))))........... (-2.32)\n"  # DP09 default, SURVEY App. C
    assert run(cli, [], stdin="gcaacgatgacatacatcgctagtcgacgc\n") == (0, want, "")
    assert run(cli, ["GCAACGAUGACAUACAUCGCUAGUCGACGC", "ignored"]) == (0, want, "")
    rc, out, err = run(cli, ["--noConv", "GCAACGATGACATACATCGCTAGTCGACGC"])
    assert (rc, out) == (0, "GCAACGATGACATACATCGCTAGTCGACGC\n....(((((.....)))))........... (-4.4)\n")
    assert err == "WARNING: stacking enthalpies not symmetric\n" * 4   # the reader's check_symmetry, like the reference
    # run from a directory without params/: the default file is cwd-relative, exactly like the reference ...
    assert run(cli, ["ACGU"], cwd="/tmp")[0] == 1
    # ... while the DNA set is linked into the binary / library (vrna_params_load_DNA_Mathews2004, src/CCJ.cc:88-90)
    rc2, out2, err2 = run(cli, ["--noConv", "GCAACGATGACATACATCGCTAGTCGACGC"], cwd="/tmp")
    assert (rc2, out2, err2) == (rc, out, err)
    if REF.exists():
        assert run(REF, ["--noConv", "GCAACGATGACATACATCGCTAGTCGACGC"], cwd="/tmp") == (rc, out, err)


@pytest.mark.gpu
def test_cli_batch_file(cli, golden_folds, tmp_path):
    """--batch-file (extension): FASTA and plain lists; per record the header, then exactly the
    reference's output for that sequence; invalid / aborting records do not stop the batch."""
    recs = [r for r in golden_folds if r["par"] == "rna_Turner04.par" and r["dangles"] == 2 and not r["extra"]
            and 8 <= len(r["seq"]) <= 70]
    ok = [r for r in recs if r["rc"] == 0][:12]
    bad = [r for r in recs if r["rc"] != 0][:2]
    use = ok[:6] + bad + ok[6:]
    fa = tmp_path / "batch.fa"
    lines = []
    for x, r in enumerate(use):
        s = r["seq"]
        lines += [f">rec{x}", s[: len(s) // 2].lower(), s[len(s) // 2:]]      # wrapped, lower case
    lines += [">broken", "ACGNNU"]
    fa.write_text("\n".join(lines) + "\n")
    rc, out, err = run(cli, ["-P", str(ROOT / "params" / "rna_Turner04.par"), "--batch-file", str(fa)])
    want = "".join(f">rec{x}\n" + r["stdout"] for x, r in enumerate(use))
    want += ">broken\nSequence contains character N that is not G,C,A,U, or T.\n"
    assert out == want
    assert err == "".join(r["stderr"] for r in use)
    assert rc == 1
    plain = tmp_path / "batch.txt"
    plain.write_text("\n".join(r["seq"] for r in ok) + "\n")
    rc, out, err = run(cli, ["-P", str(ROOT / "params" / "rna_Turner04.par"), "--batch-file", str(plain)])
    assert (rc, out, err) == (0, "".join(r["stdout"] for r in ok), "")


REFMAIN = ROOT / "oracle" / "_ref" / "CCJ_refmain_b200"


def test_reference_main_compiles_against_the_shells():
    """The reference's own src/CCJ.cc, unchanged, compiled against ccj_b200/csrc/*.hh and linked to libccj_b200.so
    (oracle/Makefile target refmain; only where /root/reference exists -- the binary travels to the GPU box)."""
    if not Path("/root/reference/src/CCJ.cc").exists():
        pytest.skip("no reference sources here")
    from ccj_b200 import build
    build.build_library()
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "refmain"], check=True, stdout=subprocess.DEVNULL)
    assert REFMAIN.exists()
    assert run(REFMAIN, ["-V"]) == (0, "CCJ 1.0\n", "")


@pytest.mark.gpu
def test_reference_main_on_the_gpu_library(golden_folds):
    """That binary folds on the GPU and prints what the reference prints."""
    if not REFMAIN.exists():
        pytest.skip("oracle/_ref/CCJ_refmain_b200 did not travel")
    recs = [r for r in golden_folds if 20 <= len(r["seq"]) <= 60][:16] + [r for r in golden_folds if r["rc"] != 0][:3]
    for r in recs:
        args = ["-P", str(ROOT / "params" / r["par"]), "-d", str(r["dangles"])] + r["extra"] + [r["seq"]]
        assert run(REFMAIN, args) == (r["rc"], r["stdout"], r["stderr"]), r["seq"]
    rc, out, err = run(REFMAIN, ["--noConv", "GCAACGATGACATACATCGCTAGTCGACGC"], cwd="/tmp")
    assert (rc, out) == (0, "GCAACGATGACATACATCGCTAGTCGACGC\n....(((((.....)))))........... (-4.4)\n")


@pytest.mark.gpu
def test_cli_takes_all_gpus_for_a_sequence_beyond_one_gpu(cli, golden_folds):
    """W_final::ccj() deals the rows of the gap tables to every GPU of the box when a sequence does not fit one
    (ccj_shard_fold); CCJ_FORCE_SHARD=1 takes that path for sequences of any length.  Needs >= 2 GPUs."""
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    recs = [r for r in golden_folds if 30 <= len(r["seq"]) <= 60 and r["par"] == "rna_Turner04.par" and not r["extra"]][:5]
    recs += [r for r in golden_folds if r["rc"] != 0][:1]
    env = dict(os.environ, CCJ_FORCE_SHARD="1")
    env.pop("NCCL_DEBUG", None)   # the box may export NCCL_DEBUG=VERSION; NCCL's own lines go to stderr (never stdout)
    for r in recs:
        args = [str(cli), "-P", str(ROOT / "params" / r["par"]), "-d", str(r["dangles"]), r["seq"]]
        p = subprocess.run(args, capture_output=True, text=True, cwd=str(ROOT), env=env)
        assert (p.returncode, p.stdout, p.stderr.replace(str(cli), "CCJ")) == (r["rc"], r["stdout"], r["stderr"]), r["seq"]
