"""Generates the golden fixtures of tests/golden/ by running the UNMODIFIED reference built by
oracle/Makefile (oracle/_ref/CCJ, oracle/_ref/ccj_ref_dump).  Run in the build container only:

    python tests/golden/make_golden.py folds      # (rc, stdout, stderr) per sequence  -> folds.json
    python tests/golden/make_golden.py hashes     # per-table FNV hashes               -> table_hashes.json
    python tests/golden/make_golden.py long       # n=100/150/200 benchmark inputs     -> folds_long.json
    python tests/golden/make_golden.py params     # scaled vrna_param_t dump           -> params_*.txt.gz
    python tests/golden/make_golden.py config4    # first 16 config-4 sequences (150 nt) -> folds_config4.json
    python tests/golden/make_golden.py config2    # first 64 config-2 sequences (100 nt) -> folds_config2.json
    python tests/golden/make_golden.py probe      # public class surface probe + struct layouts -> probe_*.txt.gz, vrna_layout.txt
    python tests/golden/make_golden.py wrap_real  # a real int16 wrap at n=87 with tripled stacking energies -> wrap_real.json
    python tests/golden/make_golden.py wrap       # int16 negative wrap (src/matrices.hh:188-191) at n=213 -> wrap213.json
    python tests/golden/make_golden.py big        # n>213: oracle/_ref/ccj_oracle hashes -> table_hashes_big.json
                                                  # (the reference aborts there; ~1 h per sequence on one core)

Sequence generators are the ones BASELINE.json / SURVEY.md 8d name for each config.
"""
import gzip
import json
import random
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = ROOT / "oracle" / "_ref" / "CCJ"
DUMP = ROOT / "oracle" / "_ref" / "ccj_ref_dump"
PARAMS = ROOT / "params"

H60 = "AAUAGGCGCAGCAUACACGGUCGAGCUGCGCCAAUAACAAUACGACCGUGAUAAAUAAAA"
K60 = "AUAGGACGCAAGGCUCGAAGCGUCCAAUAAUCCGUGCAACGAGCCAAGCACGGAUAAAAA"


# (name, parameter file, dangles, sequence) of the class-surface probes
PROBES = [("k60_t04_d2", "rna_Turner04.par", 2, K60),
          ("r38_dp09_d1", "rna_DirksPierce09.par", 1, "GGGAAACGCUCUAGCGUUUCCCAAAGAGCAAAUCGAUCA"),
          ("h44_t04_d0", "rna_Turner04.par", 0, H60[:44])]


def designed(n):
    unit = H60 + "AAUAAUAAUA" + K60 + "AAUAAUAAUA"
    return (unit * (n // len(unit) + 1))[:n]


def rand_seq(seed, n):
    rng = random.Random(seed)
    return "".join(rng.choice("ACGU") for _ in range(n))


def wrap213():
    """A 103-bp poly-G/poly-C hairpin inside the left arm of a gapped region: PK(1,211,213,213) = PK(1,1,213,213) +
    WP(2,211) falls below -32768 dcal/mol, which Matrix4D::set narrows to a positive int16 (src/matrices.hh:188-191).
    213 nt is the longest input the reference accepts, and the shortest that reaches the wrap."""
    return "G" + "G" * 103 + "GAAA" + "C" * 103 + "A" + "C"


def strong_stack_par(path):
    """rna_Turner04.par with the '# stack' free energies tripled: a 40-bp helix then reaches -39 000 dcal/mol, below
    the int16 range of Matrix4D, at a length the reference folds in seconds.  Written to `path`, returns it."""
    out, inside = [], False
    for ln in (PARAMS / "rna_Turner04.par").read_text().splitlines():
        if ln.startswith("#"):
            inside = ln.strip() == "# stack"
        elif inside and ln.strip() and not ln.lstrip().startswith("/*"):
            ln = " ".join(str(3 * int(tok)) if tok.lstrip("-").isdigit() else tok for tok in ln.split())
        out.append(ln)
    Path(path).write_text("\n".join(out) + "\n")
    return path


def wrap87():
    """40-bp poly-G/poly-C hairpin inside the left arm of a gapped region; with strong_stack_par PK(1,85,87,87) is about
    -38 000 dcal/mol in exact arithmetic and wraps to a positive int16 in the reference (src/matrices.hh:188-191)."""
    return "G" + "G" * 40 + "GAAA" + "C" * 40 + "A" + "C"


def seed200():
    random.seed(200)
    return "".join(random.choice("ACGU") for _ in range(200))


def run_ref(seq, par="rna_Turner04.par", dangles=2, extra=()):
    p = subprocess.run([str(REF), "-P", str(PARAMS / par), "-d", str(dangles), *extra, seq],
                       capture_output=True, text=True, cwd=str(ROOT))
    return {"seq": seq, "par": par, "dangles": dangles, "extra": list(extra), "rc": p.returncode,
            "stdout": p.stdout, "stderr": p.stderr}


def run_hash(seq, par="rna_Turner04.par", dangles=2):
    p = subprocess.run([str(DUMP), "hash", str(PARAMS / par), str(dangles), seq], capture_output=True, text=True,
                       cwd=str(ROOT))
    assert p.returncode == 0, p.stderr
    tabs = {}
    for line in p.stdout.splitlines()[1:]:
        name, cnt, agg, h = line.split()
        tabs[name] = [int(cnt), int(agg), h]
    return {"seq": seq, "par": par, "dangles": dangles, "tables": tabs}


def fold_jobs():
    jobs = []
    fixed = ["GCAACGAUGACAUACAUCGCUAGUCGACGC", H60, K60, "UUAGUUGUGCCGCAGCGAAGUAGUGCUUGAAAUAUGCGAC",
             "CCCUAAGUAGGAGCGUAUGCGCCCAGUAACCAAUGCCUGUUGAGAUGCCAGACGCGUAAC", "UAACGGUAGUACUAUCCAGCUCACGAGC",
             "ACGU", "A", "AC", "ACG", "GGGGAAAACCCC", "GGGAAACCC",
             "CAAAACAUAGAAACCAUCAAUAGACAGGUCAUAAUCGGUCCACCGGAUCAUUGGUGCAUAGAGCCUGGGCGUUAACGCCC"]
    for s in fixed:
        jobs.append((s, "rna_Turner04.par", 2, ()))
    for s in fixed[:6]:
        jobs.append((s, "rna_DirksPierce09.par", 2, ()))
        jobs.append((s, "rna_Turner04.par", 1, ()))
        jobs.append((s, "rna_Turner04.par", 0, ()))
        jobs.append((s, "rna_Turner04.par", 2, ("--noGU",)))
    for seed in range(240):
        rng = random.Random(7000 + seed)
        n = rng.randint(1, 56)
        par = rng.choice(["rna_Turner04.par"] * 5 + ["rna_DirksPierce09.par"])
        d = rng.choice([2, 2, 2, 2, 1, 0])
        jobs.append((rand_seq(90000 + seed, n), par, d, ()))
    return jobs


def main():
    what = sys.argv[1]
    if what == "folds":
        with ThreadPoolExecutor(8) as ex:
            out = list(ex.map(lambda j: run_ref(j[0], j[1], j[2], j[3]), fold_jobs()))
        (HERE / "folds.json").write_text(json.dumps(out, indent=0))
        print(len(out), "folds;", sum(r["rc"] != 0 for r in out), "with rc!=0;",
              sum("Should not" in r["stdout"] for r in out), "with 'Should not be here!'")
    elif what == "hashes":
        seqs = [("GCAACGAUGACAUACAUCGCUAGUCGACGC", "rna_Turner04.par", 2), (H60, "rna_Turner04.par", 2),
                (K60, "rna_Turner04.par", 2), (H60, "rna_DirksPierce09.par", 2), (K60, "rna_Turner04.par", 1),
                (K60, "rna_Turner04.par", 0)]
        seqs += [(rand_seq(500 + x, 20 + 3 * x), "rna_Turner04.par", 2) for x in range(12)]
        seqs += [(rand_seq(1000, 100)[:80], "rna_Turner04.par", 2)]
        with ThreadPoolExecutor(8) as ex:
            out = list(ex.map(lambda j: run_hash(*j), seqs))
        (HERE / "table_hashes.json").write_text(json.dumps(out, indent=0))
        print(len(out), "hash sets")
    elif what == "long":
        jobs = [(rand_seq(1000 + x, 100), "rna_Turner04.par", 2, ()) for x in range(14)]
        jobs += [(designed(100), "rna_Turner04.par", 2, ())]
        jobs += [(rand_seq(20000 + x, 150), "rna_Turner04.par", 2, ()) for x in range(4)]
        jobs += [(designed(150), "rna_Turner04.par", 2, ())]
        jobs += [(seed200(), "rna_Turner04.par", 2, ()), (designed(200), "rna_Turner04.par", 2, ())]
        jobs = jobs[::-1]  # longest first
        with ThreadPoolExecutor(8) as ex:
            out = list(ex.map(lambda j: run_ref(j[0], j[1], j[2], j[3]), jobs))
        (HERE / "folds_long.json").write_text(json.dumps(out[::-1], indent=0))
        print(len(out), "long folds")
    elif what in ("config4", "config2"):
        # parity on the first 16 config-4 and the first 64 config-2 sequences
        workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
        if what == "config4":
            jobs = [(rand_seq(20000 + x, 150), "rna_Turner04.par", 2, ()) for x in range(16)]
        else:
            jobs = [(rand_seq(1000 + x, 100), "rna_Turner04.par", 2, ()) for x in range(64)]
        with ThreadPoolExecutor(workers) as ex:
            out = list(ex.map(lambda j: run_ref(j[0], j[1], j[2], j[3]), jobs))
        (HERE / f"folds_{what}.json").write_text(json.dumps(out, indent=0))
        print(len(out), what, "folds;", sum(r["rc"] != 0 for r in out), "with rc!=0")
    elif what == "probe":
        # tests/shell/probe_body.inc run against the UNMODIFIED reference classes (ccj_ref_dump probe)
        for name, par, d, seq in PROBES:
            p = subprocess.run([str(DUMP), "probe", str(PARAMS / par), str(d), seq], capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
            with gzip.open(HERE / f"probe_{name}.txt.gz", "wt") as f:
                f.write(p.stdout)
        p = subprocess.run([str(ROOT / "oracle" / "_ref" / "layout_probe_ref")], capture_output=True, text=True)
        assert p.returncode == 0
        (HERE / "vrna_layout.txt").write_text(p.stdout)
        print("probes dumped")
    elif what == "wrap":
        seq = wrap213()
        with ThreadPoolExecutor(2) as ex:
            fa = ex.submit(run_ref, seq)
            fb = ex.submit(run_hash, seq)
            out = {"fold": fa.result(), "hashes": fb.result()}
        (HERE / "wrap213.json").write_text(json.dumps(out, indent=0))
        print("wrap213:", out["fold"]["rc"], out["fold"]["stdout"][-40:], out["hashes"]["tables"]["PK"])
    elif what == "wrap_real":
        # a REAL wrap at a length the reference folds in seconds: stacking energies tripled (custom -P file)
        import tempfile
        par = strong_stack_par(Path(tempfile.gettempdir()) / "ccj_strong_stack.par")
        seq = wrap87()
        p = subprocess.run([str(REF), "-P", str(par), seq], capture_output=True, text=True, cwd=str(ROOT))
        h = subprocess.run([str(DUMP), "hash", str(par), "2", seq], capture_output=True, text=True, cwd=str(ROOT))
        assert h.returncode == 0, h.stderr
        tabs = {}
        for line in h.stdout.splitlines()[1:]:
            name, cnt, agg, hh = line.split()
            tabs[name] = [int(cnt), int(agg), hh]
        out = {"seq": seq, "rc": p.returncode, "stdout": p.stdout, "stderr": p.stderr, "tables": tabs}
        (HERE / "wrap_real.json").write_text(json.dumps(out, indent=0))
        print("wrap_real:", p.returncode, p.stdout[-60:], tabs["PK"], tabs["PfromL"])
    elif what == "big":
        # n > 213: the reference asserts (src/matrices.hh:159-160), so the pinned CPU restatement is the checker
        sys.path.insert(0, str(ROOT))
        from oracle import oracle as orc
        seqs = [rand_seq(216, 216), designed(214)]
        def one(seq):
            tabs, w = orc.oracle_hashes(seq, "rna_Turner04.par", 2)
            return {"seq": seq, "par": "rna_Turner04.par", "dangles": 2, "tables": tabs, "W": w,
                    "source": "oracle/ccj_oracle.cc (CPU restatement; reference aborts for n >= 214)"}
        with ThreadPoolExecutor(2) as ex:
            out = list(ex.map(one, seqs))
        (HERE / "table_hashes_big.json").write_text(json.dumps(out, indent=0))
        print(len(out), "big hash sets")
    elif what == "params":
        for par in ["rna_Turner04.par", "rna_DirksPierce09.par"]:
            p = subprocess.run([str(DUMP), "params", str(PARAMS / par), "2"], capture_output=True, text=True)
            assert p.returncode == 0
            with gzip.open(HERE / f"params_{par[:-4]}.txt.gz", "wt") as f:
                f.write(p.stdout)
        # what the reference holds for sections a file omits: a file with no section, and one with two sections only
        import tempfile
        sys.path.insert(0, str(ROOT))
        sys.path.insert(0, str(ROOT / "tests"))
        from test_layout_params import STUB_PAR, partial_par
        for name, text in (("defaults", STUB_PAR), ("partial", partial_par())):
            with tempfile.NamedTemporaryFile("w", suffix=".par") as tf:
                tf.write(text)
                tf.flush()
                p = subprocess.run([str(DUMP), "params", tf.name, "2"], capture_output=True, text=True)
                assert p.returncode == 0
                with gzip.open(HERE / f"params_{name}.txt.gz", "wt") as f:
                    f.write(p.stdout)
        print("params dumped")


if __name__ == "__main__":
    main()
