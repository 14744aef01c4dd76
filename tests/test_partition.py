"""CPU tests of the multi-GPU host logic (world size 2 and 3, gloo): row blocks, ghost maps and exchange plans of
hala_b200/partition.py against a numpy restatement, and the whole exchange protocol end to end — halo exchange over
torch.distributed following the plan, local SpMV with the CPU oracle, assembled y == single-process y, bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def numpy_plan(indx_global, N, P, rank):
    """Plain numpy restatement of the contract in SURVEY.md §8(e)."""
    lo, hi = (rank * N) // P, ((rank + 1) * N) // P
    c = indx_global.astype(np.int64)
    off = (c < lo) | (c >= hi)
    ghosts = np.unique(c[off])
    local = c - lo
    local[off] = (hi - lo) + np.searchsorted(ghosts, c[off])
    bounds = np.array([(r * N) // P for r in range(P + 1)])
    owners = np.searchsorted(bounds, ghosts, side="right") - 1
    return lo, hi, local.astype(np.int32), ghosts, owners


def _worker(rank, world, port, name, n, results):
    import torch
    import torch.distributed as dist
    from hala_b200 import devgen, matgen as mg, partition as pt
    from oracle import binding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = mg.grid_rows(name, n)
        lo, hi = pt.block_range(N, world, rank)
        tp, ti, tv = devgen.stencil_slab(name, n, lo, hi, device="cpu")
        # structure of the slab is bit-identical to the numpy generator's rows
        hp, hi_, hv = mg.GENERATORS[name](n, row_lo=lo, row_hi=hi)
        assert np.array_equal(tp.numpy(), hp) and np.array_equal(ti.numpy(), hi_) and np.array_equal(tv.numpy(), hv)
        local, ghosts = pt.build_ghost_map(ti, lo, hi)
        nlo, nhi, nlocal, nghosts, nowners = numpy_plan(hi_, N, world, rank)
        assert (lo, hi) == (nlo, nhi)
        assert np.array_equal(local.numpy(), nlocal) and np.array_equal(ghosts.numpy(), nghosts)          # bit-exact maps
        assert np.array_equal(pt.owner_of(ghosts, N, world).numpy(), nowners)
        plan = pt.exchange_plan(ghosts, N, world, rank, lo)
        # invariants: ascending neighbours, counts add up, send indices are owned rows
        assert plan["neigh"] == sorted(plan["neigh"]) and rank not in plan["neigh"]
        assert sum(plan["recv_count"]) == ghosts.numel()
        s = plan["send_idx"].numpy()
        assert s.size == sum(plan["send_count"]) and (s.size == 0 or (s.min() >= 0 and s.max() < hi - lo))
        # send lists == what the numpy plan of each neighbour asks of me
        off = 0
        for q, cnt in zip(plan["neigh"], plan["send_count"]):
            qp, qi, qv = mg.GENERATORS[name](n, row_lo=pt.block_range(N, world, q)[0], row_hi=pt.block_range(N, world, q)[1])
            _, _, _, qghosts, qowners = numpy_plan(qi, N, world, q)
            assert np.array_equal(s[off:off + cnt], (qghosts[qowners == rank] - lo).astype(np.int32))
            off += cnt
        # end-to-end: exchange the halo of a probe vector following the plan, local SpMV, compare with the global product
        xg = mg.probe_x(N)
        x_ext = np.zeros(hi - lo + ghosts.numel())
        x_ext[:hi - lo] = xg[lo:hi]
        reqs, soff, roff, bufs = [], 0, 0, []
        for q, sc, rc in zip(plan["neigh"], plan["send_count"], plan["recv_count"]):
            if sc:
                reqs.append(dist.isend(torch.from_numpy(x_ext[s[soff:soff + sc]].copy()), q))
            if rc:
                buf = torch.empty(rc, dtype=torch.float64)
                bufs.append((roff, rc, buf))
                reqs.append(dist.irecv(buf, q))
            soff += sc; roff += rc
        for r in reqs:
            r.wait()
        for ro, rc, buf in bufs:
            x_ext[hi - lo + ro: hi - lo + ro + rc] = buf.numpy()
        assert np.array_equal(x_ext[hi - lo:], xg[ghosts.numpy()])
        orc = binding.oracle()
        y_local = orc.spmv(hp, local.numpy(), hv, x_ext, ncols=x_ext.size)
        gp, gi, gv = mg.GENERATORS[name](n)
        y_global = orc.spmv(gp, gi, gv, xg)
        assert np.array_equal(y_local, y_global[lo:hi])      # same entries in the same order: bit-identical
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name,n", [(2, "lap3d7", 8), (2, "lap3d27", 6), (3, "lap2d", 11), (2, "convdiff7", 5)])
def test_partition_protocol_gloo(world, name, n):
    import torch.multiprocessing as mp
    from oracle import binding
    binding.build(with_ref=False)
    port = 29500 + (os.getpid() + hash((world, name, n))) % 2000
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, name, n, results), nprocs=world, join=True)
    assert dict(results) == {r: "ok" for r in range(world)}


def test_block_ranges_cover_disjointly():
    from hala_b200 import partition as pt
    for N, P in ((134217728, 8), (1000003, 7), (5, 8), (16777216, 4)):
        edges = [pt.block_range(N, P, r) for r in range(P)]
        assert edges[0][0] == 0 and edges[-1][1] == N
        assert all(edges[r][1] == edges[r + 1][0] for r in range(P - 1))
