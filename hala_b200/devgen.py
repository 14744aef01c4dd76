"""Device-side twin of hala_b200.matgen for the structured-grid workloads (synthetic-input plumbing, PyTorch on the GPU).

The 512^3 7-point problem (BASELINE configs[2]) has 938 M non-zeros: building it with numpy on the host and pushing it
through PCIe would take minutes, so the bench builds each rank's row slab directly in HBM with the SAME integer
arithmetic as matgen._stencil (mask -> row counts -> exclusive scan -> masked select), which makes pntr/indx/vals
bit-identical to the numpy generator (tests/test_gpu_parity.py::test_device_generator_matches_host).
Nothing here is on the measured path: the arrays are handed to libhalab200 by raw pointer."""
import numpy as np
import torch

from . import matgen

_TORCH_DT = {"f32": torch.float32, "f64": torch.float64, "c32": torch.complex64, "c64": torch.complex128}


def stencil_slab(name, n, row_lo, row_hi, dtype="f64", device="cuda", delta=0.5, chunk_rows=1 << 24):
    """Rows [row_lo, row_hi) of the n^dims stencil `name` with GLOBAL column indices, built on `device`.
    Returns (pntr int32[rows+1], indx int32[nnz], vals[nnz]) as torch tensors."""
    dims = 2 if name == "lap2d" else 3
    full = name == "lap3d27"
    offs = matgen._offsets(dims, full)
    shape = (n,) * dims
    strides = [n ** (dims - 1 - d) for d in range(dims)]
    if name == "convdiff7":
        offv = [(-1.0 - delta) if sum(o) < 0 else (-1.0 + delta) for o in offs]
    else:
        offv = [-1.0] * len(offs)
    diag = {"lap2d": 4.0, "lap3d7": 6.0, "lap3d27": 26.0, "convdiff7": 6.0, "helmholtz7": 5.75 + 0.5j}[name]
    tdt = _TORCH_DT[dtype]
    S = len(offs)
    val_row = torch.tensor([diag if all(o == 0 for o in off) else offv[s] for s, off in enumerate(offs)], dtype=tdt, device=device)
    delta_row = torch.tensor([sum(off[d] * strides[d] for d in range(dims)) for off in offs], dtype=torch.int64, device=device)
    counts, cols_out, vals_out = [], [], []
    for lo in range(row_lo, row_hi, chunk_rows):
        hi = min(lo + chunk_rows, row_hi)
        rows = torch.arange(lo, hi, dtype=torch.int64, device=device)
        coords, rem = [], rows
        for d in range(dims):
            coords.append(rem // strides[d])
            rem = rem % strides[d]
        mask = torch.ones((hi - lo, S), dtype=torch.bool, device=device)
        for s, off in enumerate(offs):
            for d in range(dims):
                if off[d] != 0:
                    c = coords[d] + off[d]
                    mask[:, s] &= (c >= 0) & (c < n)
        cols = (rows[:, None] + delta_row[None, :]).to(torch.int32)
        counts.append(mask.sum(dim=1, dtype=torch.int64))
        cols_out.append(cols[mask])
        vals_out.append(val_row[None, :].expand(hi - lo, S)[mask])
        del rows, coords, rem, mask, cols
    cnt = torch.cat(counts)
    pntr = torch.zeros(row_hi - row_lo + 1, dtype=torch.int64, device=device)
    torch.cumsum(cnt, dim=0, out=pntr[1:])
    assert int(pntr[-1]) < 2 ** 31
    return pntr.to(torch.int32), torch.cat(cols_out), torch.cat(vals_out)


class torch_view:
    """Lets a torch CUDA tensor stand where hala_b200.engine expects a gpu_vector (non-owning: the tensor owns the memory)."""

    def __init__(self, engine, tensor):
        import ctypes
        assert tensor.is_cuda and tensor.is_contiguous()
        self.engine, self.tensor = engine, tensor
        self.dtype = np.dtype({torch.int32: np.int32, torch.float32: np.float32, torch.float64: np.float64,
                               torch.complex64: np.complex64, torch.complex128: np.complex128}[tensor.dtype])
        self.ptr = ctypes.c_void_p(tensor.data_ptr())
        self.num = tensor.numel()

    def size(self):
        return self.num

    def unload(self):
        return self.tensor.cpu().numpy()

    def offset(self, elements):
        import ctypes
        return ctypes.c_void_p(self.ptr.value + elements * self.dtype.itemsize)
