"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two CPU checkers:

* ``Restatement`` -> oracle/_build/liboracle.so, the plain-C restatement in oracle/restatement/.
* ``Reference``   -> oracle/_ref/libqkdref.so, the unmodified reference sources compiled in place
  (oracle/Makefile, oracle/ref_capi.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
this module. The product package (qkd_ldpc_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "_build" / "liboracle.so"
REF_SO = HERE / "_ref" / "libqkdref.so"
REF_MAIN = HERE / "_ref" / "QKD_LDPC_ref"
REFERENCE_ROOT = Path(os.environ.get("QKD_REFERENCE_ROOT", "/root/reference"))

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the restatement (always) and, when /root/reference is present, the reference itself."""
    target = "all" if ref else "oracle"
    subprocess.run(["make", "-C", str(HERE), target], check=True, capture_output=True)


class _Graph(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("m", C.c_int32),
        ("e", C.c_int32),
        ("row_ptr", C.c_void_p),
        ("col_idx", C.c_void_p),
        ("col_ptr", C.c_void_p),
        ("row_idx", C.c_void_p),
    ]


class _SP(C.Structure):
    _fields_ = [("iterations_num", C.c_uint64), ("syndromes_match", C.c_int32)]


class _LDPC(C.Structure):
    _fields_ = [("sp_res", _SP), ("keys_match", C.c_int32)]


class Graph:
    """Flattened H (CSR over checks + CSC over bits), list order preserved. 0-based."""

    def __init__(self, n, m, row_ptr, col_idx, col_ptr, row_idx, is_regular=None, max_bit_w=None, max_check_w=None):
        self.n, self.m = int(n), int(m)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        self.col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
        self.col_ptr = np.ascontiguousarray(col_ptr, dtype=np.int32)
        self.row_idx = np.ascontiguousarray(row_idx, dtype=np.int32)
        self.e = int(self.col_idx.size)
        assert self.row_idx.size == self.e and self.row_ptr.size == self.m + 1 and self.col_ptr.size == self.n + 1
        bw, cw = np.diff(self.col_ptr), np.diff(self.row_ptr)
        self.is_regular = bool((bw == bw[0]).all() and (cw == cw[0]).all()) if is_regular is None else bool(is_regular)
        self.max_bit_w = int(bw.max()) if max_bit_w is None else int(max_bit_w)
        self.max_check_w = int(cw.max()) if max_check_w is None else int(max_check_w)
        self._c = _Graph(self.n, self.m, self.e, self.row_ptr.ctypes.data, self.col_idx.ctypes.data,
                         self.col_ptr.ctypes.data, self.row_idx.ctypes.data)

    @property
    def rate(self) -> float:
        return 1.0 - self.m / self.n

    @staticmethod
    def from_check_lists(n, check_lists):
        """Build both halves from per-check sorted bit lists (what the dense loader produces,
        ref: src/array_and_matrix_operations.cpp:4-47)."""
        m = len(check_lists)
        row_ptr = np.zeros(m + 1, np.int32)
        row_ptr[1:] = np.cumsum([len(r) for r in check_lists])
        col_idx = np.concatenate([np.asarray(r, np.int32) for r in check_lists]) if m else np.zeros(0, np.int32)
        bit_lists = [[] for _ in range(n)]
        for j, r in enumerate(check_lists):
            for b in r:
                bit_lists[b].append(j)
        col_ptr = np.zeros(n + 1, np.int32)
        col_ptr[1:] = np.cumsum([len(c) for c in bit_lists])
        row_idx = np.concatenate([np.asarray(c, np.int32) for c in bit_lists]) if n else np.zeros(0, np.int32)
        return Graph(n, m, row_ptr, col_idx, col_ptr, row_idx)

    @staticmethod
    def from_dense(h):
        h = np.asarray(h)
        return Graph.from_check_lists(h.shape[1], [np.flatnonzero(r) for r in h])


def parse_dense(path) -> Graph:
    """ref: src/array_and_matrix_operations.cpp:295-421 (values 0/1, equal row lengths, no empty row/column)."""
    rows = [[int(t) for t in line.split()] for line in Path(path).read_text().splitlines()]
    if not rows:
        raise RuntimeError(f"File is empty or cannot be read properly: {path}")
    h = np.array(rows)
    if not np.isin(h, (0, 1)).all():
        raise RuntimeError("Parity check matrix can only take values 0 or 1.")
    if (h.sum(0) <= 0).any() or (h.sum(1) <= 0).any():
        raise RuntimeError("row/column weight cannot be equal to or less than zero")
    return Graph.from_dense(h)


def parse_alist(path) -> Graph:
    """ref: src/array_and_matrix_operations.cpp:109-292. 1-based in the file, 0-based in memory; each list is
    read up to the node's weight (trailing zero padding ignored); the two halves are NOT cross-checked."""
    lines = [[int(t) for t in line.split()] for line in Path(path).read_text().splitlines()]
    if len(lines) < 4 or len(lines[0]) != 2 or len(lines[1]) != 2:
        raise RuntimeError(f"File format does not match the alist format: {path}")
    n, m = lines[0]
    max_bw, max_cw = lines[1]
    bw, cw = lines[2], lines[3]
    if n != len(bw) or m != len(cw) or len(lines) < 4 + n + m:
        raise RuntimeError(f"Insufficient or inconsistent data in the file: {path}")
    for i in range(n):
        if sum(1 for v in lines[4 + i] if v != 0) != bw[i]:
            raise RuntimeError(f"Number of non-zero elements in line {4 + i + 1} does not match the weight")
    for j in range(m):
        if sum(1 for v in lines[4 + n + j] if v != 0) != cw[j]:
            raise RuntimeError(f"Number of non-zero elements in line {4 + n + j + 1} does not match the weight")
    col_ptr = np.zeros(n + 1, np.int32)
    col_ptr[1:] = np.cumsum(bw)
    row_ptr = np.zeros(m + 1, np.int32)
    row_ptr[1:] = np.cumsum(cw)
    row_idx = np.array([v - 1 for i in range(n) for v in lines[4 + i][: bw[i]]], np.int32)
    col_idx = np.array([v - 1 for j in range(m) for v in lines[4 + n + j][: cw[j]]], np.int32)
    regular = all(w == bw[0] for w in bw) and all(w == cw[0] for w in cw)
    return Graph(n, m, row_ptr, col_idx, col_ptr, row_idx, is_regular=regular, max_bit_w=max_bw, max_check_w=max_cw)


def fnv1a64_bits(bits) -> str:
    """64-bit FNV-1a over one byte per bit -- the hash the survey's golden table uses."""
    h = 0xCBF29CE484222325
    for b in np.asarray(bits, dtype=np.uint8).tobytes():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


class Restatement:
    """The plain-C restatement (oracle/restatement/sp_oracle.c)."""

    F32_DIVIDE, F32_LEAVE_ONE_OUT = 0, 1

    def __init__(self, so_path: Path = ORACLE_SO):
        if not Path(so_path).exists():
            build(ref=False)
        L = self.lib = C.CDLL(str(so_path))
        gp = C.POINTER(_Graph)
        L.orc_syndrome.argtypes = [gp, _i32p, _i32p]
        L.orc_sum_product_f64.argtypes = [gp, _f64p, _i32p, C.c_uint64, C.c_int, C.c_double, _i32p]
        L.orc_sum_product_f64.restype = _SP
        L.orc_sum_product_f64_trace.argtypes = [gp, _f64p, _i32p, C.c_uint64, C.c_int, C.c_double, C.c_uint64, _f64p, _f64p,
                                                _i32p, _i32p, _f64p, _i32p]
        L.orc_sum_product_f64_trace.restype = _SP
        L.orc_sum_product_f32.argtypes = [gp, _f32p, _i32p, C.c_uint64, C.c_int, C.c_float, C.c_int, _i32p]
        L.orc_sum_product_f32.restype = _SP
        L.orc_qkd_ldpc.argtypes = [gp, _i32p, _i32p, C.c_double, C.c_uint64, C.c_int, C.c_double, C.c_int, C.c_int,
                                   _i32p, _i32p]
        L.orc_qkd_ldpc.restype = _LDPC
        L.orc_trial_seeds.argtypes = [C.c_uint64, C.c_uint64, _u64p]
        L.orc_generate.argtypes = [C.c_uint64, C.c_uint64, C.c_double, _i32p, _i32p]
        L.orc_generate.restype = C.c_double
        L.orc_run_trials.argtypes = [gp, C.c_double, _u64p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_double,
                                     C.c_int, C.c_int, _u64p, C.c_void_p]
        L.orc_run_trials.restype = C.c_int
        L.orc_qber_range.argtypes = [C.c_double, _f64p, C.c_size_t, _f64p, C.c_size_t]
        L.orc_qber_range.restype = C.c_int
        L.orc_point_stats.argtypes = [_u64p, C.c_uint64, C.c_uint64, _f64p]

    def syndrome(self, g: Graph, bits):
        out = np.zeros(g.m, np.int32)
        self.lib.orc_syndrome(C.byref(g._c), np.ascontiguousarray(bits, np.int32), out)
        return out

    def sum_product(self, g: Graph, llr, syndrome, max_it=100, thr=100.0, enable_thr=True, precision=64,
                    f32_form=F32_LEAVE_ONE_OUT):
        out = np.zeros(g.n, np.int32)
        syn = np.ascontiguousarray(syndrome, np.int32)
        if precision == 64:
            r = self.lib.orc_sum_product_f64(C.byref(g._c), np.ascontiguousarray(llr, np.float64), syn, max_it,
                                             int(enable_thr), thr, out)
        else:
            r = self.lib.orc_sum_product_f32(C.byref(g._c), np.ascontiguousarray(llr, np.float32), syn, max_it,
                                             int(enable_thr), thr, f32_form, out)
        return int(r.iterations_num), bool(r.syndromes_match), out

    def sum_product_trace(self, g: Graph, llr, syndrome, capacity, max_it=100, thr=100.0, enable_thr=True):
        """fp64 decode with the reference's TRACE_SUM_PRODUCT intermediates for the first `capacity` iterations."""
        cap = min(int(capacity), int(max_it))
        rows = max(cap, 1)
        e, m = np.zeros((rows, g.e), np.float64), np.zeros((rows, g.e), np.float64)
        tot = np.zeros((rows, g.n), np.float64)
        z, s = np.zeros((rows, g.n), np.int32), np.zeros((rows, g.m), np.int32)
        out = np.zeros(g.n, np.int32)
        r = self.lib.orc_sum_product_f64_trace(C.byref(g._c), np.ascontiguousarray(llr, np.float64),
                                               np.ascontiguousarray(syndrome, np.int32), max_it, int(enable_thr), thr, cap,
                                               e.reshape(-1), tot.reshape(-1), z.reshape(-1), s.reshape(-1), m.reshape(-1), out)
        return dict(E=e[:cap], L=tot[:cap], z=z[:cap], s=s[:cap], M=m[:cap], bits=out, iterations=int(r.iterations_num),
                    result=int(r.syndromes_match))

    def qkd_ldpc(self, g: Graph, alice, bob, qber, max_it=100, thr=100.0, enable_thr=True, precision=64,
                 f32_form=F32_LEAVE_ONE_OUT):
        syn = np.zeros(g.m, np.int32)
        dec = np.zeros(g.n, np.int32)
        r = self.lib.orc_qkd_ldpc(C.byref(g._c), np.ascontiguousarray(alice, np.int32),
                                  np.ascontiguousarray(bob, np.int32), qber, max_it, int(enable_thr), thr, precision,
                                  f32_form, syn, dec)
        return int(r.sp_res.iterations_num), bool(r.sp_res.syndromes_match), bool(r.keys_match), syn, dec

    def trial_seeds(self, simulation_seed: int, count: int):
        out = np.zeros(count, np.uint64)
        self.lib.orc_trial_seeds(simulation_seed, count, out)
        return out

    def generate(self, seed: int, n: int, qber: float):
        a = np.zeros(n, np.int32)
        b = np.zeros(n, np.int32)
        exact = self.lib.orc_generate(int(seed), n, qber, a, b)
        return a, b, float(exact)

    def run_trials(self, g: Graph, qber, seeds, threads=1, max_it=100, thr=100.0, enable_thr=True, precision=64,
                   f32_form=F32_LEAVE_ONE_OUT, want_decoded=False):
        seeds = np.ascontiguousarray(seeds, np.uint64)
        out3 = np.zeros(3 * seeds.size, np.uint64)
        dec = np.zeros((seeds.size, g.n), np.int32) if want_decoded else None
        rc = self.lib.orc_run_trials(C.byref(g._c), qber, seeds, seeds.size, threads, max_it, int(enable_thr), thr,
                                     precision, f32_form, out3, dec.ctypes.data if want_decoded else None)
        if rc != 0:
            raise RuntimeError(f"Key size '{g.n}' is too small for QBER.")
        out3 = out3.reshape(-1, 3)
        return (out3, dec) if want_decoded else out3

    def qber_range(self, code_rate, params):
        p = np.ascontiguousarray(np.asarray(params, np.float64).reshape(-1, 4))
        out = np.zeros(4096, np.float64)
        k = self.lib.orc_qber_range(code_rate, p.ravel(), p.shape[0], out, out.size)
        if k < 0:
            raise RuntimeError("An error occurred when generating a QBER range based on code rate.")
        return out[:k].copy()

    def point_stats(self, out3, max_it=100):
        out3 = np.ascontiguousarray(out3, np.uint64).reshape(-1, 3)
        s = np.zeros(6, np.float64)
        self.lib.orc_point_stats(out3.ravel(), out3.shape[0], max_it, s)
        return dict(mean=s[0], std_dev=s[1], min=s[2], max=s[3], ratio_sp=s[4], ratio_ldpc=s[5])


class Reference:
    """The unmodified reference, compiled in place (oracle/_ref/libqkdref.so)."""

    def __init__(self, so_path: Path = REF_SO, max_it=100, thr=100.0, enable_thr=True, threads=1):
        if not Path(so_path).exists():
            if (REFERENCE_ROOT / "src").exists():
                build(ref=True)
            else:
                raise FileNotFoundError(f"{so_path} is not built and {REFERENCE_ROOT} is absent")
        L = self.lib = C.CDLL(str(so_path))
        L.ref_last_error.restype = C.c_char_p
        L.ref_set_cfg.argtypes = [C.c_uint64, C.c_int, C.c_double, C.c_uint64]
        L.ref_matrix_load.argtypes = [C.c_char_p, C.c_int]
        L.ref_matrix_load.restype = C.c_void_p
        L.ref_matrix_free.argtypes = [C.c_void_p]
        L.ref_matrix_info.argtypes = [C.c_void_p, _u64p]
        L.ref_matrix_export.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, _i32p]
        L.ref_trial_seeds.argtypes = [C.c_uint64, C.c_uint64, _u64p]
        L.ref_prng_raw.argtypes = [C.c_uint64, C.c_uint64, _u64p]
        L.ref_generate.argtypes = [C.c_uint64, C.c_uint64, C.c_double, _i32p, _i32p]
        L.ref_generate.restype = C.c_double
        L.ref_syndrome.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int]
        L.ref_sum_product.argtypes = [C.c_void_p, _f64p, _i32p, C.c_uint64, C.c_double, _i32p, _u64p, C.c_int]
        L.ref_qkd_ldpc.argtypes = [C.c_void_p, _i32p, _i32p, C.c_double, _u64p, C.c_int]
        L.ref_run_trial.argtypes = [C.c_void_p, C.c_double, C.c_uint64, _u64p, C.POINTER(C.c_double)]
        L.ref_run_trial.restype = C.c_int
        L.ref_run_trials.argtypes = [C.c_void_p, C.c_double, _u64p, C.c_uint64, C.c_uint64, _u64p]
        L.ref_run_trials.restype = C.c_int
        L.ref_qber_range.argtypes = [C.c_double, _f64p, C.c_uint64, _f64p, C.c_uint64]
        L.ref_qber_range.restype = C.c_int
        self.set_cfg(max_it, thr, enable_thr, threads)

    def set_cfg(self, max_it=100, thr=100.0, enable_thr=True, threads=1):
        self.max_it, self.thr, self.enable_thr = max_it, thr, enable_thr
        self.lib.ref_set_cfg(max_it, int(enable_thr), thr, threads)

    def error(self) -> str:
        return self.lib.ref_last_error().decode()

    def load(self, path, dense=False):
        h = self.lib.ref_matrix_load(str(path).encode(), int(dense))
        if not h:
            raise RuntimeError(self.error())
        return h

    def free(self, h):
        self.lib.ref_matrix_free(h)

    def graph(self, h) -> Graph:
        info = np.zeros(7, np.uint64)
        self.lib.ref_matrix_info(h, info)
        n, m, mbw, mcw, reg, eb, ec = (int(v) for v in info)
        assert eb == ec, "the two adjacency halves disagree on the edge count"
        row_ptr, col_idx = np.zeros(m + 1, np.int32), np.zeros(ec, np.int32)
        col_ptr, row_idx = np.zeros(n + 1, np.int32), np.zeros(eb, np.int32)
        self.lib.ref_matrix_export(h, row_ptr, col_idx, col_ptr, row_idx)
        return Graph(n, m, row_ptr, col_idx, col_ptr, row_idx, is_regular=bool(reg), max_bit_w=mbw, max_check_w=mcw)

    def trial_seeds(self, simulation_seed, count):
        out = np.zeros(count, np.uint64)
        self.lib.ref_trial_seeds(simulation_seed, count, out)
        return out

    def prng_raw(self, seed, count):
        out = np.zeros(count, np.uint64)
        self.lib.ref_prng_raw(seed, count, out)
        return out

    def generate(self, seed, n, qber):
        a, b = np.zeros(n, np.int32), np.zeros(n, np.int32)
        exact = self.lib.ref_generate(int(seed), n, qber, a, b)
        return a, b, float(exact)

    def syndrome(self, h, bits, m, variant=-1):
        out = np.zeros(m, np.int32)
        self.lib.ref_syndrome(h, np.ascontiguousarray(bits, np.int32), out, variant)
        return out

    def sum_product(self, h, n, llr, syndrome, variant=-1):
        out = np.zeros(n, np.int32)
        r = np.zeros(2, np.uint64)
        self.lib.ref_sum_product(h, np.ascontiguousarray(llr, np.float64), np.ascontiguousarray(syndrome, np.int32),
                                 self.max_it, self.thr, out, r, variant)
        return int(r[0]), bool(r[1]), out

    def qkd_ldpc(self, h, alice, bob, qber, variant=-1):
        r = np.zeros(3, np.uint64)
        self.lib.ref_qkd_ldpc(h, np.ascontiguousarray(alice, np.int32), np.ascontiguousarray(bob, np.int32), qber, r,
                              variant)
        return int(r[0]), bool(r[1]), bool(r[2])

    def run_trial(self, h, qber, seed):
        r = np.zeros(3, np.uint64)
        exact = C.c_double()
        if self.lib.ref_run_trial(h, qber, int(seed), r, C.byref(exact)) != 0:
            raise RuntimeError(self.error())
        return int(r[0]), bool(r[1]), bool(r[2]), exact.value

    def run_trials(self, h, qber, seeds, threads=1):
        seeds = np.ascontiguousarray(seeds, np.uint64)
        out3 = np.zeros(3 * seeds.size, np.uint64)
        if self.lib.ref_run_trials(h, qber, seeds, seeds.size, threads, out3) != 0:
            raise RuntimeError(self.error())
        return out3.reshape(-1, 3)

    def qber_range(self, code_rate, params):
        p = np.ascontiguousarray(np.asarray(params, np.float64).reshape(-1, 4))
        out = np.zeros(4096, np.float64)
        k = self.lib.ref_qber_range(code_rate, p.ravel(), p.shape[0], out, out.size)
        if k < 0:
            raise RuntimeError(self.error())
        return out[:k].copy()
