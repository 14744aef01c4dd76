"""1-D row-block partition, ghost-index maps and exchange plans for the row-partitioned solvers (host logic of SURVEY.md §8e).

Written with device-agnostic torch ops, so the same code builds the maps for CUDA tensors on the GPU box and for CPU tensors
in the gloo tests (tests/test_partition.py).  The reference has no multi-device path; the contract is our own and is pinned
by a numpy restatement in the tests plus invariants.

    block_range(N, P, r)          rank r owns rows [floor(rN/P), floor((r+1)N/P))
    build_ghost_map(...)          columns renumbered to [owned | ghosts]; ghosts = sorted unique off-block global columns
    exchange_plan(...)            who sends which owned entries to whom (lists exchanged over torch.distributed)
"""
import torch


def block_range(N, P, r):
    return (r * N) // P, ((r + 1) * N) // P


def owner_of(cols, N, P):
    """Rank owning global column c under block_range: the largest r with floor(rN/P) <= c."""
    r = (cols * P + P - 1) // N                     # first guess, then correct the integer-rounding cases
    r = torch.clamp(r, 0, P - 1)
    lo = (r * N) // P
    r = torch.where(lo > cols, r - 1, r)
    hi = ((r + 1) * N) // P
    r = torch.where(cols >= hi, r + 1, r)
    return r


def build_ghost_map(indx_global, row_lo, row_hi):
    """indx_global: int32 tensor of GLOBAL column indices of the local rows.
    Returns (indx_local int32, ghosts int64 sorted unique).  Owned column c -> c - row_lo; ghost g -> n_owned + rank of g in ghosts."""
    n_owned = row_hi - row_lo
    c = indx_global.to(torch.int64)
    off = (c < row_lo) | (c >= row_hi)
    ghosts = torch.unique(c[off])                   # sorted
    local = c - row_lo
    if ghosts.numel() > 0:
        pos = torch.searchsorted(ghosts, c[off])
        local[off] = n_owned + pos
    return local.to(torch.int32), ghosts


def exchange_plan(ghosts, N, P, rank, row_lo, group=None, device=None):
    """Per-neighbour counts and the send-index list of this rank.
    Ghosts are sorted by global index, hence already grouped by owner in ascending rank order: neighbour k's block of
    received entries is contiguous in [n_owned + recv_off[k], ...).  Every rank tells every owner which of its columns it
    needs (all_gather of the padded request lists — the lists are short: the halo); the owner turns them into local indices.
    Returns dict(neigh, send_count, recv_count, send_idx)  with send_idx an int32 tensor on `device`."""
    import torch.distributed as dist
    device = device if device is not None else ghosts.device
    owners = owner_of(ghosts, N, P) if ghosts.numel() else ghosts
    need_from = torch.bincount(owners, minlength=P) if ghosts.numel() else torch.zeros(P, dtype=torch.int64, device=ghosts.device)
    need_from = need_from.to(torch.int64)
    # 1. everybody learns the full request-count matrix: counts[q][r] = how many columns rank q needs from rank r
    cnt_list = [torch.zeros(P, dtype=torch.int64, device=need_from.device) for _ in range(P)]
    dist.all_gather(cnt_list, need_from, group=group)
    counts = torch.stack(cnt_list).cpu()
    # 2. everybody publishes its (padded) sorted ghost list; owners pick their segment out of each requester's list
    max_g = int(counts.sum(dim=1).max())
    padded = torch.full((max(max_g, 1),), -1, dtype=torch.int64, device=ghosts.device)
    padded[:ghosts.numel()] = ghosts
    all_lists = [torch.empty_like(padded) for _ in range(P)]
    dist.all_gather(all_lists, padded, group=group)
    neigh, send_count, recv_count, send_parts = [], [], [], []
    for q in range(P):
        if q == rank:
            continue
        s, r = int(counts[q][rank]), int(counts[rank][q])
        if s == 0 and r == 0:
            continue
        neigh.append(q); send_count.append(s); recv_count.append(r)
        if s > 0:
            start = int(counts[q][:rank].sum())     # q's ghosts are sorted, so its requests to lower ranks come first
            send_parts.append((all_lists[q][start:start + s] - row_lo).to(torch.int32))
    send_idx = torch.cat(send_parts) if send_parts else torch.zeros(0, dtype=torch.int32, device=ghosts.device)
    return {"neigh": neigh, "send_count": send_count, "recv_count": recv_count, "send_idx": send_idx.to(device)}
