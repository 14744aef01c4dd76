"""Deterministic synthetic CSR matrices, right-hand sides and probe vectors (host side, numpy).

These are the workloads BASELINE.json names; the reference ships no generators, only
tiny hand-written CSR fixtures (tests/sparse_tests.hpp:166-191), so the definitions here ARE the contract:
the same arrays are handed to the reference cpu_engine path (oracle) and to the CUDA path, which makes the
CSR structure bit-identical on both sides by construction. All matrices: int32 0-based `pntr`/`indx`, rows
sorted by column, diagonal present — the layout `hala::sparse_gemv` takes (sparse/hala_sparse_blas.hpp:79-96).

    lap2d(n)            C1  2-D 5-point Laplacian, index i*n+j, diag 4, neighbours -1, Dirichlet truncation
    lap3d27(n)          C2  3-D 27-point, index (k*n+j)*n+i, diag 26, 26 neighbours -1
    lap3d7(n)           C3  3-D 7-point, diag 6, neighbours -1
    convdiff7(n, d)     C4  7-point convection-diffusion: lower neighbours -1-d, upper -1+d (nonsymmetric)
    helmholtz7(n)       C5a complex 7-point, diag 5.75+0.5i, neighbours -1
    powerlaw(N, ...)    C5b irregular row lengths l_i = clamp(floor(lmin*(1-u)^(-1/(a-1))), 1, lmax)
"""
import numpy as np

DTYPES = {"f32": np.float32, "f64": np.float64, "c32": np.complex64, "c64": np.complex128}
DTYPE_CODE = {"f32": 0, "f64": 1, "c32": 2, "c64": 3}

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed, idx):
    """splitmix64 finaliser of (seed + (idx+1)*golden); idx may be an array. Returns uint64."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (np.asarray(idx, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hashed_uniform(seed, n, lo=-1.0, hi=1.0, offset=0):
    """n doubles uniform in [lo, hi): 53 high bits of splitmix64(seed, offset+i)."""
    h = splitmix64(seed, np.arange(offset, offset + n, dtype=np.uint64))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


def _stencil(shape, offsets, values, diag_value, dtype, row_lo=0, row_hi=None):
    """Generic structured-grid stencil on `shape` (slowest .. fastest), lexicographic row index,
    Dirichlet truncation. `offsets` are index-space tuples sorted so that column order is ascending;
    rows [row_lo, row_hi) only (global column indices) — used by the row partitioner tests."""
    dims = len(shape)
    N = int(np.prod(shape))
    row_hi = N if row_hi is None else row_hi
    rows = np.arange(row_lo, row_hi, dtype=np.int64)
    coords = np.unravel_index(rows, shape)
    strides = [int(np.prod(shape[d + 1:])) for d in range(dims)]
    S = len(offsets)
    mask = np.empty((rows.size, S), dtype=bool)
    cols = np.empty((rows.size, S), dtype=np.int32)
    vals = np.empty((rows.size, S), dtype=dtype)
    for s, off in enumerate(offsets):
        ok = np.ones(rows.size, dtype=bool)
        delta = 0
        for d in range(dims):
            if off[d] != 0:
                c = coords[d] + off[d]
                ok &= (c >= 0) & (c < shape[d])
            delta += off[d] * strides[d]
        mask[:, s] = ok
        cols[:, s] = (rows + delta).astype(np.int32)
        vals[:, s] = diag_value if all(o == 0 for o in off) else values[s]
    pntr = np.zeros(rows.size + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=pntr[1:])
    assert pntr[-1] < 2**31
    return pntr.astype(np.int32), cols[mask], vals[mask]


def _offsets(dims, full):
    rng = (-1, 0, 1)
    if dims == 2:
        offs = [(a, b) for a in rng for b in rng]
    else:
        offs = [(a, b, c) for a in rng for b in rng for c in rng]
    if not full:
        offs = [o for o in offs if sum(abs(v) for v in o) <= 1]
    return offs  # already lexicographic == ascending column order


def lap2d(n, dtype="f64", **kw):
    offs = _offsets(2, False)
    return _stencil((n, n), offs, [-1.0] * len(offs), 4.0, DTYPES[dtype], **kw)


def lap3d7(n, dtype="f64", **kw):
    offs = _offsets(3, False)
    return _stencil((n, n, n), offs, [-1.0] * len(offs), 6.0, DTYPES[dtype], **kw)


def lap3d27(n, dtype="f64", **kw):
    offs = _offsets(3, True)
    return _stencil((n, n, n), offs, [-1.0] * len(offs), 26.0, DTYPES[dtype], **kw)


def convdiff7(n, delta=0.5, dtype="f64", **kw):
    offs = _offsets(3, False)
    vals = [(-1.0 - delta) if sum(o) < 0 else (-1.0 + delta) for o in offs]
    return _stencil((n, n, n), offs, vals, 6.0, DTYPES[dtype], **kw)


def helmholtz7(n, dtype="c64", **kw):
    offs = _offsets(3, False)
    return _stencil((n, n, n), offs, [-1.0] * len(offs), 5.75 + 0.5j, DTYPES[dtype], **kw)


def powerlaw(N=1 << 22, alpha=2.5, lmin=4, lmax=65536, seed=42, dtype="f64"):
    """Irregular CSR that stresses the row-length binning: heavy-tailed row lengths, hashed distinct columns
    (diagonal always present, columns sorted), off-diagonals -(0.5+0.5 v) e^{i theta} (theta = 0 for real
    dtypes), diagonal = 1 + sum|off| (strictly diagonally dominant, so GMRES converges quickly)."""
    dt = DTYPES[dtype]
    u = hashed_uniform(seed, N, 0.0, 1.0)
    length = np.floor(lmin * (1.0 - u) ** (-1.0 / (alpha - 1.0)))
    length = np.clip(length, 1, min(lmax, N)).astype(np.int64)
    pntr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(length, out=pntr[1:])
    nnz = int(pntr[-1])
    assert nnz < 2**31
    row_of = np.repeat(np.arange(N, dtype=np.int64), length)
    slot = np.arange(nnz, dtype=np.int64) - pntr[row_of]
    # distinct columns per row: a stratified hash — slot s of a row of length L draws from the s-th of L
    # equal strata of [0, N), so columns are distinct and already ascending; slot 0's stratum is replaced
    # by nothing special: the diagonal is forced in by overwriting the slot whose stratum contains i.
    L = length[row_of]
    lo = (slot * N) // L
    hi = ((slot + 1) * N) // L
    h = splitmix64(seed + 1, np.arange(nnz, dtype=np.uint64))
    cols = lo + (h % np.maximum(hi - lo, 1).astype(np.uint64)).astype(np.int64)
    diag_slot = (row_of * L) // N           # stratum that contains column == row
    # make sure the stratum really contains the row index (integer rounding): fix by search of neighbours
    for adj in (0, 1, -1):
        s = np.clip(diag_slot + adj, 0, L - 1)
        inside = ((s * N) // L <= row_of) & (row_of < ((s + 1) * N) // L)
        diag_slot = np.where(inside, s, diag_slot)
    is_diag = slot == diag_slot
    cols = np.where(is_diag, row_of, cols)
    v = (splitmix64(seed + 2, np.arange(nnz, dtype=np.uint64)) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    mag = 0.5 + 0.5 * v
    if np.issubdtype(dt, np.complexfloating):
        theta = 2.0 * np.pi * (splitmix64(seed + 3, np.arange(nnz, dtype=np.uint64)) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
        off = -mag * np.exp(1j * theta)
    else:
        off = -mag
    absoff = np.where(is_diag, 0.0, mag)
    rowsum = np.add.reduceat(absoff, pntr[:-1])
    vals = np.where(is_diag, 1.0 + rowsum[row_of], off).astype(dt)
    return pntr.astype(np.int32), cols.astype(np.int32), vals


def rhs(N, dtype="f64", seed=11):
    """b with ||b||_2 = 1: ones/sqrt(N) (real part); complex dtypes add a hashed imaginary part then normalise."""
    dt = DTYPES[dtype]
    if np.issubdtype(dt, np.complexfloating):
        b = np.ones(N) + 1j * hashed_uniform(seed, N, -0.5, 0.5)
        return (b / np.linalg.norm(b)).astype(dt)
    return np.full(N, 1.0 / np.sqrt(N), dtype=dt)


def probe_x(N, dtype="f64", seed=7):
    """SpMV probe vector: hashed uniform(-1,1) (not smooth, so per-entry relative error is meaningful)."""
    dt = DTYPES[dtype]
    if np.issubdtype(dt, np.complexfloating):
        return (hashed_uniform(seed, N) + 1j * hashed_uniform(seed + 1000, N)).astype(dt)
    return hashed_uniform(seed, N).astype(dt)


GENERATORS = {"lap2d": lap2d, "lap3d7": lap3d7, "lap3d27": lap3d27, "convdiff7": convdiff7,
              "helmholtz7": helmholtz7}


def grid_rows(name, n):
    return n * n if name == "lap2d" else n * n * n


def spmv_bytes(N, nnz, itemsize):
    """Algorithmic bytes of one SpMV (beta = 0): SURVEY.md §8(d)."""
    return nnz * (itemsize + 4) + (N + 1) * 4 + 2 * N * itemsize


def cg_iter_bytes(N, nnz, itemsize):
    """Algorithmic bytes of one fused CG iteration (3-kernel schedule): SURVEY.md §8(d)."""
    return nnz * (itemsize + 4) + (N + 1) * 4 + 11 * N * itemsize


def gmres_iter_bytes(N, nnz, itemsize, restart):
    """Algorithmic bytes of one GMRES(m) inner iteration averaged over a restart cycle (basis size k = 1..m): SpMV, multi-dot
    (k columns + the vector once per chunk of 16 columns), multi-axpy + norm (k columns, r read and written), normalise-and-append
    (read r, write the new column): DESIGN.md §4."""
    ks = range(1, restart + 1)
    per_k = [(k + (k + 15) // 16) + (k + 2) + 2 for k in ks]
    return spmv_bytes(N, nnz, itemsize) + int(sum(per_k) / len(per_k) * N * itemsize)


# ---------------------------------------------------------------- triangular / ILU test inputs
F1_CASES = (("lap3d27", 6), ("convdiff7", 7), ("lap2d", 13))      # (generator, size) of the recorded trsv / ILU parity cases


def perturbed(name, n, dtype="f64", seed=77, eps=0.1):
    """Generator `name` with a deterministic hashed perturbation of every value (makes the stencils nonsymmetric and, for the
    complex types, genuinely complex) — the matrices the triangular-solve and ILU parity cases are built from."""
    p, i, v = GENERATORS[name](n, dtype=dtype)
    return p, i, (v + eps * probe_x(v.size, dtype, seed=seed)).astype(v.dtype)


def split_triangle(pntr, indx, vals, uplo):
    """One-triangle CSR (diagonal included) of a CSR with sorted rows: the layout the reference's cpu_triangular_matrix expects
    (diagonal LAST in each row for 'L', FIRST for 'U'; sparse/hala_sparse_utils.hpp:283-335)."""
    n = pntr.size - 1
    rows = np.repeat(np.arange(n), np.diff(pntr))
    keep = indx <= rows if uplo in "Ll" else indx >= rows
    tp = np.concatenate([[0], np.cumsum(np.bincount(rows[keep], minlength=n))]).astype(np.int32)
    return tp, np.ascontiguousarray(indx[keep]), np.ascontiguousarray(vals[keep])
