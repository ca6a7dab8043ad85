"""Parity-check matrix containers and file formats.

`Matrix` is the flattened form of the reference's `H_matrix` (src/array_and_matrix_operations.hpp:16-27): both
adjacency halves, stored order preserved. Readers follow the reference's loaders:
  read_alist  <- read_sparse_alist_matrix (src/array_and_matrix_operations.cpp:109-292)
  read_dense  <- read_dense_matrix        (src/array_and_matrix_operations.cpp:295-421)
The shipped codes travel as compact .npz files under data/codes (the GPU box has no reference checkout);
`materialize()` writes them back out as alist / dense text files for the C++ loaders.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np

DATA = Path(__file__).resolve().parent.parent / "data"
CODES = DATA / "codes"
GENERATED = DATA / "_generated"
NORTH_STAR = "n10240_m5231_cw3_seed666"


@dataclass
class Matrix:
    n: int
    m: int
    row_ptr: np.ndarray
    col_idx: np.ndarray
    col_ptr: np.ndarray
    row_idx: np.ndarray
    is_regular: bool = False
    max_bit_w: int = 0
    max_check_w: int = 0
    name: str = ""

    @property
    def e(self) -> int:
        return int(self.col_idx.size)

    @property
    def rate(self) -> float:
        return 1.0 - self.m / self.n

    @staticmethod
    def from_lists(n, check_lists, bit_lists, name="", max_bit_w=None, max_check_w=None):
        m = len(check_lists)
        cw = np.array([len(r) for r in check_lists], np.int64)
        bw = np.array([len(c) for c in bit_lists], np.int64)
        row_ptr = np.zeros(m + 1, np.int32)
        row_ptr[1:] = np.cumsum(cw)
        col_ptr = np.zeros(n + 1, np.int32)
        col_ptr[1:] = np.cumsum(bw)
        col_idx = np.concatenate([np.asarray(r, np.int32) for r in check_lists]) if m else np.zeros(0, np.int32)
        row_idx = np.concatenate([np.asarray(c, np.int32) for c in bit_lists]) if n else np.zeros(0, np.int32)
        regular = bool((bw == bw[0]).all() and (cw == cw[0]).all())
        return Matrix(n, m, row_ptr, col_idx.astype(np.int32), col_ptr, row_idx.astype(np.int32), regular,
                      int(bw.max()) if max_bit_w is None else int(max_bit_w),
                      int(cw.max()) if max_check_w is None else int(max_check_w), name)

    @staticmethod
    def from_dense(h, name=""):
        h = np.asarray(h)
        return Matrix.from_lists(h.shape[1], [np.flatnonzero(r) for r in h], [np.flatnonzero(c) for c in h.T], name)

    def to_dense(self) -> np.ndarray:
        h = np.zeros((self.m, self.n), np.uint8)
        for j in range(self.m):
            h[j, self.col_idx[self.row_ptr[j]:self.row_ptr[j + 1]]] = 1
        return h


def load_npz(name_or_path) -> Matrix:
    p = Path(name_or_path)
    if not p.exists():
        p = CODES / f"{name_or_path}.npz"
    z = np.load(p)
    n, m = int(z["n"]), int(z["m"])
    row_ptr = np.zeros(m + 1, np.int32)
    row_ptr[1:] = np.cumsum(z["check_w"].astype(np.int64))
    col_ptr = np.zeros(n + 1, np.int32)
    col_ptr[1:] = np.cumsum(z["bit_w"].astype(np.int64))
    return Matrix(n, m, row_ptr, z["col_idx"].astype(np.int32), col_ptr, z["row_idx"].astype(np.int32), bool(z["is_regular"]),
                  int(z["max_bit_w"]), int(z["max_check_w"]), str(z["source"]))


def read_alist(path) -> Matrix:
    lines = [[int(t) for t in ln.split()] for ln in Path(path).read_text().splitlines()]
    if not lines:
        raise RuntimeError(f"File is empty or cannot be read properly: {path}")
    if len(lines) < 4:
        raise RuntimeError(f"Insufficient data in the file: {path}")
    if len(lines[0]) != 2 or len(lines[1]) != 2:
        raise RuntimeError(f"File format does not match the alist format: {path}")
    (n, m), (max_bw, max_cw), bw, cw = lines[0], lines[1], lines[2], lines[3]
    if len(lines) < 4 + len(bw) + len(cw):
        raise RuntimeError(f"Insufficient data in the file: {path}")
    if n != len(bw):
        raise RuntimeError(f"Number of columns '{n}' is not the same as the length of the third line '{len(bw)}'. File: {path}")
    if m != len(cw):
        raise RuntimeError(f"Number of rows '{m}' is not the same as the length of the fourth line '{len(cw)}'. File: {path}")
    for i, w in enumerate(bw):
        nz = sum(1 for v in lines[4 + i] if v != 0)
        if nz != w:
            raise RuntimeError(f"Number of non-zero elements '{nz}' in the line '{4 + i + 1}' does not match the weight in the "
                               f"third line '{w}'. File: {path}")
    for j, w in enumerate(cw):
        nz = sum(1 for v in lines[4 + n + j] if v != 0)
        if nz != w:
            raise RuntimeError(f"Number of non-zero elements '{nz}' in the line '{4 + n + j + 1}' does not match the weight in "
                               f"the fourth line '{w}'. File: {path}")
    bit_lists = [[v - 1 for v in lines[4 + i][:bw[i]]] for i in range(n)]
    check_lists = [[v - 1 for v in lines[4 + n + j][:cw[j]]] for j in range(m)]
    mat = Matrix.from_lists(n, check_lists, bit_lists, Path(path).name, max_bw, max_cw)
    return mat


def read_dense(path) -> Matrix:
    rows = [[int(t) for t in ln.split()] for ln in Path(path).read_text().splitlines()]
    if not rows:
        raise RuntimeError(f"File is empty or cannot be read properly: {path}")
    if any(v not in (0, 1) for r in rows for v in r):
        raise RuntimeError("Parity check matrix can only take values 0 or 1.")
    if any(len(r) != len(rows[0]) for r in rows):
        raise RuntimeError(f"Different lengths of rows in a matrix. File: {path}")
    h = np.array(rows)
    for i, w in enumerate(h.sum(0)):
        if w <= 0:
            raise RuntimeError(f"Column '{i + 1}' weight cannot be equal to or less than zero. File: {path}")
    for j, w in enumerate(h.sum(1)):
        if w <= 0:
            raise RuntimeError(f"Row '{j + 1}' weight cannot be equal to or less than zero. File: {path}")
    return Matrix.from_dense(h, Path(path).name)


def write_alist(mat: Matrix, path) -> None:
    """Standard alist: header, weights, then 1-based lists zero-padded to the maximum weight."""
    bw, cw = np.diff(mat.col_ptr), np.diff(mat.row_ptr)
    mb, mc = max(int(bw.max()), mat.max_bit_w), max(int(cw.max()), mat.max_check_w)
    out = [f"{mat.n} {mat.m}", f"{mb} {mc}", " ".join(map(str, bw)), " ".join(map(str, cw))]
    for i in range(mat.n):
        lst = (mat.row_idx[mat.col_ptr[i]:mat.col_ptr[i + 1]] + 1).tolist()
        out.append(" ".join(map(str, lst + [0] * (mb - len(lst)))))
    for j in range(mat.m):
        lst = (mat.col_idx[mat.row_ptr[j]:mat.row_ptr[j + 1]] + 1).tolist()
        out.append(" ".join(map(str, lst + [0] * (mc - len(lst)))))
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    Path(path).write_text("\n".join(out) + "\n")


def write_dense(mat: Matrix, path) -> None:
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    Path(path).write_text("\n".join(" ".join(map(str, r)) for r in mat.to_dense().tolist()))


def materialize() -> dict:
    """Writes every data/codes/*.npz back out as the text file the reference ships (same file names)."""
    out = {}
    for p in sorted(CODES.glob("*.npz")):
        z = np.load(p)
        mat = load_npz(p)
        sub = GENERATED / ("dense_matrices" if bool(z["dense"]) else "alist_sparse_matrices")
        dst = sub / str(z["source"])
        if not dst.exists():
            (write_dense if bool(z["dense"]) else write_alist)(mat, dst)
        out[p.stem] = dst
    return out


def save_npz(mat: Matrix, name: str) -> Path:
    """Writes a matrix in the format load_npz reads (data/codes/<name>.npz)."""
    idx_t = np.uint16 if max(mat.n, mat.m) < 65536 else np.uint32
    p = CODES / f"{name}.npz"
    np.savez_compressed(p, n=mat.n, m=mat.m, is_regular=mat.is_regular, max_bit_w=mat.max_bit_w, max_check_w=mat.max_check_w,
                        check_w=np.diff(mat.row_ptr).astype(np.uint8), col_idx=mat.col_idx.astype(idx_t),
                        bit_w=np.diff(mat.col_ptr).astype(np.uint8), row_idx=mat.row_idx.astype(idx_t), source=mat.name, dense=False)
    return p


def peg_code(n: int, m: int, dv: int = 3, seed: int = 666, bfs_limit: int = 4096) -> Matrix:
    """A seeded PEG code of column weight `dv` (host/peg.cpp), cached as alist under data/_generated/peg/ -- the
    construction the shipped N=10240 code is consistent with; used for BASELINE.json configs[3] and configs[4].
    A committed copy data/codes/peg_n<n>_m<m>_cw<dv>_seed<seed>_bfs<limit>.npz (the generator's own output) is used when present."""
    import subprocess
    from . import build
    cached = CODES / f"peg_n{n}_m{m}_cw{dv}_seed{seed}_bfs{bfs_limit}.npz"
    if cached.exists():
        return load_npz(cached)
    rate = 1.0 - m / n
    path = GENERATED / "peg" / f"(N={n},M={m},R={rate:.2f},CW={dv},SEED={seed}).txt"
    if not path.exists():
        build.build_host()
        subprocess.run([str(build.SIM_PATH), "--peg", str(n), str(m), str(dv), str(seed), str(path), str(bfs_limit)], check=True)
    return read_alist_fast(path)


def read_alist_fast(path) -> Matrix:
    """read_alist without the per-line validation loops (generated files, up to millions of lines)."""
    with open(path) as f:
        n, m = map(int, f.readline().split())
        max_bw, max_cw = map(int, f.readline().split())
        bw = np.array(f.readline().split(), np.int64)
        cw = np.array(f.readline().split(), np.int64)
        rows = [np.array(f.readline().split(), np.int64) for _ in range(n)]
        cols = [np.array(f.readline().split(), np.int64) for _ in range(m)]
    row_idx = np.concatenate([r[:w] for r, w in zip(rows, bw)]) - 1
    col_idx = np.concatenate([c[:w] for c, w in zip(cols, cw)]) - 1
    col_ptr = np.zeros(n + 1, np.int32); col_ptr[1:] = np.cumsum(bw)
    row_ptr = np.zeros(m + 1, np.int32); row_ptr[1:] = np.cumsum(cw)
    regular = bool((bw == bw[0]).all() and (cw == cw[0]).all())
    return Matrix(n, m, row_ptr, col_idx.astype(np.int32), col_ptr, row_idx.astype(np.int32), regular, max_bw, max_cw, Path(path).name)


def permutation_code(n: int, m: int, dv: int = 3, seed: int = 666) -> Matrix:
    """A seeded column-weight-`dv` code from `dv` random permutations (Gallager's construction): layer l deals the bits, in
    the order of a random permutation, round-robin to the checks (start offset l * m / dv), so every check gets floor or
    ceil(n / m) bits per layer; a bit that drew the same check in two layers trades its entry with another bit's.
    O(E) in numpy -- seconds at N = 1 000 000, where the PEG construction (host/peg.cpp) takes minutes -- for the large-frame
    roofline and parity cases of BASELINE.json configs[3]. No girth conditioning: a throughput / parity workload, not a code
    design. Adjacency lists sorted ascending, as in the shipped files."""
    rng = np.random.default_rng(seed)
    check_of = np.empty((dv, n), np.int64)
    for layer in range(dv):
        perm = rng.permutation(n)
        check_of[layer, perm] = (np.arange(n) + layer * (m // dv)) % m
    for _ in range(64):
        clean = True
        for hi in range(1, dv):
            for lo in range(hi):
                for b in np.flatnonzero(check_of[lo] == check_of[hi]):  # trade the later layer's entry with another bit's
                    o = int(rng.integers(n))
                    check_of[hi, b], check_of[hi, o] = check_of[hi, o], check_of[hi, b]
                    clean = False
        if clean:
            break
    else:
        raise RuntimeError("permutation_code: could not remove the double edges")
    row_idx = np.sort(check_of, axis=0).T.reshape(-1).astype(np.int32)  # bit-major, checks ascending
    col_ptr = (np.arange(n + 1) * dv).astype(np.int32)
    bits = np.repeat(np.arange(n, dtype=np.int64), dv)
    order = np.lexsort((bits, row_idx))  # check-major, bits ascending inside a check
    col_idx = bits[order].astype(np.int32)
    cw = np.bincount(row_idx, minlength=m)
    if (cw == 0).any():
        raise RuntimeError("permutation_code: a check received no bit")
    row_ptr = np.zeros(m + 1, np.int32)
    row_ptr[1:] = np.cumsum(cw)
    regular = bool((cw == cw[0]).all())
    return Matrix(n, m, row_ptr, col_idx, col_ptr, row_idx, regular, dv, int(cw.max()), f"(N={n},M={m},CW={dv},SEED={seed},PERM)")
