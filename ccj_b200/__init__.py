"""ccj_b200 -- B200-native CCJ (Chen-Condon-Jabbari) MFE pseudoknot fold.

Python host mirror of the reference's driver (src/CCJ.cc:44-49: ``ccj(seq, energy, dangle)``) on top of
the C ABI of ``include/ccj_b200.h`` (ctypes, no torch types cross the boundary).  The compute path is
the CUDA library ``libccj_b200.so``; there is no CPU fallback -- loading fails loudly when the library
or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Iterable, List, Optional, Sequence

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libccj_b200.so"
PARAMS_DIR = _PKG.parent / "params"  # the reference's params/*.par data files, unchanged

TABLE4 = ["PK", "PL", "PR", "PM", "PO", "PfromL", "PfromR", "PfromM", "PfromMprime", "PfromO",
          "PLmloop00", "PLmloop01", "PLmloop10", "PRmloop00", "PRmloop01", "PRmloop10",
          "PMmloop00", "PMmloop01", "PMmloop10", "POmloop00", "POmloop01", "POmloop10"]
TABLE2 = ["V", "Vtype", "WM", "WMv", "WMp", "P", "WBP", "WPP", "WB", "WP"]

_PREFIX = ["", "border case: ", "border cases: ", "boder cases: ", "impossible case: ", "impossible cases: ",
           "impossbible cases: "]
_NODE = {'P': "P_P", 'k': "P_PK", 'l': "P_PL", 'r': "P_PR", 'm': "P_PM", 'o': "P_PO", 'f': "P_PfromL",
         'g': "P_PfromR", 'h': "P_PfromM", '[': "P_PfromMprime", ']': "P_PfromMdoubleprime", 'i': "P_PfromO",
         'j': "P_PLiloop", 'c': "P_PLmloop", 'e': "P_PLmloop10", 'n': "P_PLmloop01", 'a': "P_PLmloop00",
         'q': "P_PRiloop", 't': "P_PRmloop", 'u': "P_PRmloop10", '&': "P_PRmloop01", '9': "P_PRmloop00",
         'w': "P_PMiloop", 'y': "P_PMmloop", '0': "P_PMmloop10", '1': "P_PMmloop01", '8': "P_PMmloop00",
         'z': "P_POiloop", '+': "P_POmloop", '-': "P_POmloop10", '=': "P_POmloop01", '_': "P_POmloop00",
         '*': "P_WB", '^': "P_WBP", '#': "P_WP", '@': "P_WPP"}


class CCJError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ccj_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class _Result(C.Structure):
    _fields_ = [("energy_dcal", C.c_int32), ("status", C.c_int32), ("n_should_not_be_here", C.c_int32),
                ("msg_id", C.c_int32), ("aux_i", C.c_int32), ("aux_j", C.c_int32)]


RESULT_DTYPE = np.dtype([("energy_dcal", "<i4"), ("status", "<i4"), ("n_should_not_be_here", "<i4"),
                         ("msg_id", "<i4"), ("aux_i", "<i4"), ("aux_j", "<i4")])

_lib = None


def load_library(path: Optional[os.PathLike] = None) -> C.CDLL:
    """dlopen the CUDA library. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else Path(os.environ.get("CCJ_B200_LIB", LIB_PATH))   # CCJ_B200_LIB: an alternative build (experiments)
    if not p.exists():
        raise CCJError(-1, f"{p} is missing: build it with `python -m ccj_b200.build` (nvcc, sm_100a); "
                           "there is no CPU fallback")
    lib = C.CDLL(str(p))
    vp, i32, i64p = C.c_void_p, C.c_int, C.POINTER(C.c_int64)
    lib.ccj_version.restype = C.c_char_p
    lib.ccj_ctx_create.argtypes = [i32, C.POINTER(vp)]
    lib.ccj_ctx_destroy.argtypes = [vp]
    lib.ccj_ctx_destroy.restype = None
    lib.ccj_last_error.argtypes = [vp]
    lib.ccj_last_error.restype = C.c_char_p
    lib.ccj_model_load.argtypes = [vp, C.c_char_p, i32, i32]
    lib.ccj_model_load_embedded.argtypes = [vp, C.c_char_p, i32, i32]
    lib.ccj_fold_batch.argtypes = [vp, vp, i64p, i32, vp, vp, vp]
    lib.ccj_fold_batch_multi.argtypes = [C.POINTER(vp), i32, vp, i64p, i32, vp, vp, vp]
    lib.ccj_batch_prepare.argtypes = [vp, vp, i64p, i32]
    lib.ccj_batch_fill.argtypes = [vp]
    lib.ccj_batch_traceback.argtypes = [vp]
    lib.ccj_batch_fetch.argtypes = [vp, vp, vp, vp]
    lib.ccj_last_fill_ms.argtypes = [vp]
    lib.ccj_last_fill_ms.restype = C.c_float
    lib.ccj_last_traceback_ms.argtypes = [vp]
    lib.ccj_last_traceback_ms.restype = C.c_float
    lib.ccj_last_fill_launches.argtypes = [vp]
    lib.ccj_wave_capacity.argtypes = [vp, i32]
    lib.ccj_wave_capacity.restype = C.c_int64
    lib.ccj_stream.argtypes = [vp]
    lib.ccj_stream.restype = vp
    lib.ccj_export_table4.argtypes = [vp, i32, i32, vp, C.c_int64]
    lib.ccj_export_table2.argtypes = [vp, i32, i32, vp, C.c_int64]
    lib.ccj_table4_len.argtypes = [i32]
    lib.ccj_table4_len.restype = C.c_int64
    lib.ccj_table2_len.argtypes = [i32]
    lib.ccj_table2_len.restype = C.c_int64
    lib.ccj_batch_fill_profiled.argtypes = [vp, C.POINTER(C.c_float)]
    lib.ccj_count_terms.argtypes = [C.c_char_p, i32, i32, i64p]
    lib.ccj_measure_addmin_peak.argtypes = [vp, i32, C.POINTER(C.c_double)]
    i32p = C.POINTER(C.c_int32)
    lib.ccj_table4_get.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32p]
    lib.ccj_table2_get.argtypes = [vp, i32, i32, i32, i32, i32p]
    u64p = C.POINTER(C.c_uint64)
    lib.ccj_table4_hash.argtypes = [vp, i32, i32, u64p, i64p, C.POINTER(C.c_int32)]
    lib.ccj_table2_hash.argtypes = [vp, i32, i32, u64p, i64p, i64p]
    lib.ccj_model_text.argtypes = [C.c_char_p, i32, i32, C.c_char_p, C.c_char_p, C.c_size_t]
    lib.ccj_layout_index.argtypes = [i32, i32, i32, i32, i32]
    lib.ccj_layout_index.restype = C.c_int64
    if path is None:
        _lib = lib
    return lib


def default_par_file(name: str = "rna_Turner04.par") -> str:
    """Energy parameter files shipped with the package (same files as the reference's params/)."""
    return str(PARAMS_DIR / name)


def prepare_sequence(seq: str, no_conv: bool = False) -> str:
    """Upper-case and T->U exactly like src/CCJ.cc:74-75."""
    s = seq.upper()
    if not no_conv:
        s = s.replace("T", "U")
    return s


def traceback_message(msg_id: int) -> str:
    p, t = divmod(msg_id, 256)
    return f"{_PREFIX[p] if 0 <= p < len(_PREFIX) else ''}This should not have happened!, {_NODE.get(chr(t), '?')}"


def energy_text(energy_dcal: int) -> str:
    """`std::cout << double` formatting of W[n]/100.0 (src/CCJ.cc:108)."""
    return "%g" % (energy_dcal / 100.0)


@dataclass
class Fold:
    """Outcome of one sequence, with the reference binary's observable behaviour re-created."""
    sequence: str
    structure: str
    energy_dcal: int
    status: int
    n_should_not_be_here: int
    msg_id: int
    aux_i: int = 0
    aux_j: int = 0

    @property
    def energy(self) -> float:
        return self.energy_dcal / 100.0

    @property
    def returncode(self) -> int:
        return 1 if self.status == 1 else 0

    @property
    def stdout(self) -> str:
        out = "Should not be here!\n" * self.n_should_not_be_here
        if self.status == 0:
            out += f"{self.sequence}\n{self.structure} ({energy_text(self.energy_dcal)})\n"
        return out

    @property
    def stderr(self) -> str:
        if self.status == 1:
            return traceback_message(self.msg_id) + "\n"
        if self.status == 2:
            return (f"NOT GOOD RESTR INTER, i={self.aux_i}, j={self.aux_j}, best_ip={self.aux_j}, "
                    f"best_jp={self.aux_i}\n")
        return ""


class Context:
    """One GPU + one energy model (replaces the reference's process-global parameter state)."""

    def __init__(self, device: int = 0, par_file: Optional[str] = None, dangles: int = 2, no_gu: bool = False):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.ccj_ctx_create(device, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise CCJError(rc, "cannot create a CUDA context (no B200 visible?); there is no CPU fallback")
        self.device = device
        self.load_model(par_file or default_par_file(), dangles, no_gu)

    # -- lifetime -------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.ccj_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise CCJError(rc, self._lib.ccj_last_error(self._h).decode())

    # -- model ----------------------------------------------------------------------------------
    def load_model(self, par_file: str, dangles: int = 2, no_gu: bool = False) -> None:
        """`par_file` is a path, or "@dna_mathews2004" / "@rna_turner2004" for a set linked into the library."""
        if str(par_file).startswith("@"):
            self._check(self._lib.ccj_model_load_embedded(self._h, str(par_file)[1:].encode(), int(dangles),
                                                          int(bool(no_gu))))
        else:
            self._check(self._lib.ccj_model_load(self._h, str(par_file).encode(), int(dangles), int(bool(no_gu))))
        self.par_file, self.dangles, self.no_gu = str(par_file), dangles, no_gu

    # -- folding --------------------------------------------------------------------------------
    @staticmethod
    def _pack(seqs: Sequence[str]):
        blob = "".join(seqs).encode("ascii")
        offs = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=offs[1:])
        return blob, offs

    def _unpack(self, seqs, offs, res, structs) -> List[Fold]:
        out = []
        sb = structs.tobytes().decode("ascii")
        for x, s in enumerate(seqs):
            r = res[x]
            out.append(Fold(s, sb[offs[x]:offs[x + 1]], int(r["energy_dcal"]), int(r["status"]),
                            int(r["n_should_not_be_here"]), int(r["msg_id"]), int(r["aux_i"]), int(r["aux_j"])))
        return out

    def fold_batch(self, seqs: Iterable[str]) -> List[Fold]:
        """W_final(seq, dangle).ccj() for every sequence (independent; waves sized to GPU memory)."""
        seqs = list(seqs)
        if not seqs:
            return []
        blob, offs = self._pack(seqs)
        res = np.zeros(len(seqs), dtype=RESULT_DTYPE)
        structs = np.zeros(len(blob), dtype=np.uint8)
        self._check(self._lib.ccj_fold_batch(self._h, blob, offs.ctypes.data_as(C.POINTER(C.c_int64)), len(seqs),
                                             res.ctypes.data, None, structs.ctypes.data))
        return self._unpack(seqs, offs, res, structs)

    def fold(self, seq: str) -> Fold:
        return self.fold_batch([seq])[0]

    # -- split phases (measurement / table inspection) -------------------------------------------
    def prepare(self, seqs: Sequence[str]) -> None:
        self._seqs = list(seqs)
        self._blob, self._offs = self._pack(self._seqs)
        self._check(self._lib.ccj_batch_prepare(self._h, self._blob, self._offs.ctypes.data_as(C.POINTER(C.c_int64)),
                                                len(self._seqs)))

    def fill(self) -> float:
        self._check(self._lib.ccj_batch_fill(self._h))
        return float(self._lib.ccj_last_fill_ms(self._h))

    def fill_profiled(self):
        """Fill launched kernel by kernel; returns summed device ms of (K_4D, K_P, K_2D, other)."""
        ms = (C.c_float * 6)()
        self._check(self._lib.ccj_batch_fill_profiled(self._h, ms))
        return {"k4d_ms": ms[0] + ms[4] + ms[5], "k4d_split_ms": ms[0], "k4d_window_ms": ms[4], "k4d_final_ms": ms[5],
                "kP_ms": ms[1], "k2d_ms": ms[2], "other_ms": ms[3],
                "total_ms": float(self._lib.ccj_last_fill_ms(self._h))}

    def traceback(self) -> float:
        self._check(self._lib.ccj_batch_traceback(self._h))
        return float(self._lib.ccj_last_traceback_ms(self._h))

    def fetch(self) -> List[Fold]:
        res = np.zeros(len(self._seqs), dtype=RESULT_DTYPE)
        structs = np.zeros(len(self._blob), dtype=np.uint8)
        self._check(self._lib.ccj_batch_fetch(self._h, res.ctypes.data, None, structs.ctypes.data))
        return self._unpack(self._seqs, self._offs, res, structs)

    @property
    def last_fill_ms(self) -> float:
        return float(self._lib.ccj_last_fill_ms(self._h))

    @property
    def last_traceback_ms(self) -> float:
        return float(self._lib.ccj_last_traceback_ms(self._h))

    @property
    def last_fill_launches(self) -> int:
        return int(self._lib.ccj_last_fill_launches(self._h))

    def addmin_peak(self) -> dict:
        """Measured integer ceiling of this GPU in min-plus candidates per second (include/ccj_b200.h)."""
        out = {}
        for variant, name in enumerate(["int32_viaddmnmx", "int16_unpack_viaddmnmx", "int16x2_viaddmnmx"]):
            v = C.c_double()
            self._check(self._lib.ccj_measure_addmin_peak(self._h, variant, C.byref(v)))
            out[name] = v.value
        return out

    @property
    def stream_ptr(self) -> int:
        """The cudaStream_t the context launches on (for callers that time with their own CUDA events)."""
        return int(self._lib.ccj_stream(self._h))

    def wave_capacity(self, n: int) -> int:
        return int(self._lib.ccj_wave_capacity(self._h, n))

    def table4(self, seq_index: int, table) -> np.ndarray:
        """int16 values of one gap table in the canonical export order (i, j>=i, k>=j+2, l>=k)."""
        t = TABLE4.index(table) if isinstance(table, str) else int(table)
        n = len(self._seqs[seq_index])
        out = np.empty(self._lib.ccj_table4_len(n), dtype=np.int16)
        self._check(self._lib.ccj_export_table4(self._h, seq_index, t, out.ctypes.data, out.size))
        return out

    def table4_hash(self, seq_index: int, table):
        """(finite count, min, fnv hex) exactly as `ccj_ref_dump hash` prints them."""
        t = TABLE4.index(table) if isinstance(table, str) else int(table)
        h, fin, mn = C.c_uint64(), C.c_int64(), C.c_int32()
        self._check(self._lib.ccj_table4_hash(self._h, seq_index, t, C.byref(h), C.byref(fin), C.byref(mn)))
        return [fin.value, mn.value, "%016x" % h.value]

    def table2_hash(self, seq_index: int, table):
        t = TABLE2.index(table) if isinstance(table, str) else int(table)
        h, fin, sm = C.c_uint64(), C.c_int64(), C.c_int64()
        self._check(self._lib.ccj_table2_hash(self._h, seq_index, t, C.byref(h), C.byref(fin), C.byref(sm)))
        return [fin.value, sm.value, "%016x" % h.value]

    def table2(self, seq_index: int, table) -> np.ndarray:
        t = TABLE2.index(table) if isinstance(table, str) else int(table)
        n = len(self._seqs[seq_index])
        out = np.empty(self._lib.ccj_table2_len(n), dtype=np.int32)
        self._check(self._lib.ccj_export_table2(self._h, seq_index, t, out.ctypes.data, out.size))
        return out


def fold_batch_multi(contexts: Sequence["Context"], seqs: Iterable[str]) -> List[Fold]:
    """`Context.fold_batch` over several GPUs from one process: one host thread per context inside the library,
    sequences dealt dynamically (ccj_fold_batch_multi)."""
    seqs = list(seqs)
    if not seqs:
        return []
    ctx0 = contexts[0]
    blob, offs = Context._pack(seqs)
    res = np.zeros(len(seqs), dtype=RESULT_DTYPE)
    structs = np.zeros(len(blob), dtype=np.uint8)
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    rc = ctx0._lib.ccj_fold_batch_multi(arr, len(contexts), blob, offs.ctypes.data_as(C.POINTER(C.c_int64)), len(seqs),
                                        res.ctypes.data, None, structs.ctypes.data)
    ctx0._check(rc)
    return ctx0._unpack(seqs, offs, res, structs)


def model_text(par_file: str, dangles: int = 2, no_gu: bool = False) -> str:
    """Scaled energy model as text (host only, no GPU), same format as `ccj_ref_dump params`."""
    import tempfile
    lib = load_library()
    err = C.create_string_buffer(512)
    with tempfile.NamedTemporaryFile("r", suffix=".txt") as f:
        rc = lib.ccj_model_text(str(par_file).encode(), dangles, int(no_gu), f.name.encode(), err, 512)
        if rc != 0:
            raise CCJError(rc, err.value.decode())
        return f.read()


def count_terms(seq: str, no_gu: bool = False) -> dict:
    """Algorithmic work of one fold: cells, split terms, P terms, evaluated window terms,
    and the derived algorithmic bytes: 2*(split + 2*P + iloop) + 44*cells."""
    out = (C.c_int64 * 4)()
    rc = load_library().ccj_count_terms(seq.encode(), len(seq), int(no_gu), out)
    if rc != 0:
        raise CCJError(rc, "ccj_count_terms")
    d = {"cells": out[0], "split": out[1], "pterms": out[2], "iloop": out[3]}
    d["bytes_4d"] = 2 * (d["split"] + d["iloop"]) + 44 * d["cells"]
    d["bytes_P"] = 2 * 2 * d["pterms"]
    d["bytes"] = d["bytes_4d"] + d["bytes_P"]
    return d


def layout_index(n: int, i: int, j: int, k: int, l: int) -> int:
    return int(load_library().ccj_layout_index(n, i, j, k, l))


def ccj(seq: str, dangles: int = 2, par_file: Optional[str] = None, device: int = 0) -> Fold:
    """One-shot convenience mirroring src/CCJ.cc:44-49."""
    with Context(device, par_file, dangles) as ctx:
        return ctx.fold(prepare_sequence(seq))


def cells(n: int) -> int:
    """Number of 4D DP cells of a length-n fold, C(n+1,4)."""
    return (n + 1) * n * (n - 1) * (n - 2) // 24 if n >= 3 else 0
