"""Block-sharded all-pairs Chamfer / EMD matrices and the MMD / COV / 1-NNA metrics built on them.

The reference ships the two metric operators (chamfer_3D, earth_mover_distance) but no evaluation script
(SURVEY.md section 0); BASELINE.json config 4 defines the workload: 1000 generated x 1000 reference clouds of
2048 points.  Every matrix entry is the reference operator applied to one cloud pair:

    CD[i,j]  = dist1.mean() + dist2.mean()            of chamfer_3D.forward(S_i, R_j)   (dist_chamfer_3D.py:44-67)
    EMD[i,j] = earth_mover_distance(S_i, R_j, transpose=False)                          (emd.py:27-51)

The matrix is a set of independent pairs, so it shards with no data-path collective: rank k owns a contiguous
block of ROWS.  NCCL (gloo in the CPU tests) is used only to all-gather a reference set that was loaded sharded
and to all-gather the finished row blocks.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "all_gather_clouds", "pairwise_matrix", "pairwise_matrix_sharded",
           "symmetric_matrix_sharded", "mmd_cov", "one_nna", "compute_metrics"]


def shard_range(n, rank, world):
    """Contiguous balanced split of range(n): the first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_clouds(local, counts=None, group=None):
    """All-gather a cloud set [n_local, N, 3] whose rows are sharded over ranks (row counts may differ by one)."""
    rank, world = _world(group)
    if world == 1:
        return local
    if counts is None:
        c = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
        cs = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(cs, c, group=group)
        counts = [int(x.item()) for x in cs]
    mx = max(counts)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:k] for b, k in zip(bufs, counts)], dim=0)


def _batched_pairs(metric):
    """Pair function on the batched OPERATORS (chamfer_3D.forward / approxmatch + matchcost, one batch per row cloud):
    for clouds beyond the limits of the all-pairs kernels (Chamfer: 8192 points in the second set; EMD: both clouds
    resident in shared memory, ~5k points each).  Same values, without the one-CTA-per-pair kernels' speed."""
    from . import backend

    def cd(a, b):
        nb, m = b.shape[0], b.shape[1]
        out = a.new_empty((a.shape[0], nb))
        d1 = a.new_empty((nb, a.shape[1])); d2 = a.new_empty((nb, m))
        i1 = torch.empty((nb, a.shape[1]), dtype=torch.int32, device=a.device)
        i2 = torch.empty((nb, m), dtype=torch.int32, device=a.device)
        for i in range(a.shape[0]):
            if backend.chamfer.forward(a[i:i + 1].expand(nb, -1, -1).contiguous(), b, d1, d2, i1, i2) != 1:
                raise RuntimeError("chamfer_3D.forward failed")
            out[i] = d1.mean(1) + d2.mean(1)
        return out

    def emd(a, b):
        nb = b.shape[0]
        out = a.new_empty((a.shape[0], nb))
        for i in range(a.shape[0]):
            x = a[i:i + 1].expand(nb, -1, -1).contiguous()
            out[i] = backend.emd.matchcost_forward(x, b, backend.emd.approxmatch_forward(x, b)) / a.shape[1]  # emd.py:48-50
        return out

    return {"cd": cd, "emd": emd}[metric]


def _default_pair_fn(metric):
    from . import backend  # imported lazily: CPU-only tests inject their own pair function
    fast = {"cd": backend.chamfer3d_allpairs, "emd": backend.emd_allpairs}[metric]
    slow = _batched_pairs(metric)

    def fn(a, b):
        n, m = a.shape[1], b.shape[1]
        fits = m <= 8192 if metric == "cd" else 20 * n + 24 * m <= 220 * 1024   # chamfer.cu kApMaxM / emd.cu shared memory
        return fast(a, b) if fits else slow(a.contiguous(), b.contiguous())

    return fn


def pairwise_matrix(a, b, metric="cd", pair_fn=None, block=256):
    """Full [na, nb] matrix on this device, computed in row blocks of `block` clouds."""
    fn = pair_fn or _default_pair_fn(metric)
    rows = [fn(a[i:i + block].contiguous(), b) for i in range(0, a.shape[0], block)]
    return torch.cat(rows, dim=0) if rows else a.new_zeros((0, b.shape[0]))


def pairwise_matrix_sharded(a, b, metric="cd", pair_fn=None, group=None, block=256, gather=True):
    """Rows of `a` (replicated on every rank) are split over ranks; returns the full matrix on every rank
    (gather=True) or (row_lo, row_hi, local_block)."""
    rank, world = _world(group)
    lo, hi = shard_range(a.shape[0], rank, world)
    local = pairwise_matrix(a[lo:hi], b, metric, pair_fn, block)
    if not gather:
        return lo, hi, local
    if world == 1:
        return local
    counts = [shard_range(a.shape[0], r, world) for r in range(world)]
    counts = [h - l for l, h in counts]
    return all_gather_clouds(local, counts, group)


def symmetric_matrix_sharded(a, metric="cd", pair_fn=None, group=None, block=256):
    """[n, n] matrix of a set against itself when the metric is symmetric (Chamfer: dist1.mean() + dist2.mean() with
    the roles swapped is the same two terms): only the blocks on and above the diagonal are computed — half the pairs
    — and mirrored.  The upper-triangle blocks are dealt to the ranks round-robin by descending size; one all-reduce
    (every entry is written by exactly one rank, the others add zero) leaves the full matrix on every rank."""
    rank, world = _world(group)
    fn = pair_fn or _default_pair_fn(metric)
    n = a.shape[0]
    out = a.new_zeros((n, n))
    starts = list(range(0, n, block))
    jobs = [(i, j) for i in starts for j in starts if j >= i]
    jobs.sort(key=lambda ij: -(min(block, n - ij[0]) * min(block, n - ij[1])))
    for k, (i, j) in enumerate(jobs):
        if k % world != rank:
            continue
        blk = fn(a[i:i + block].contiguous(), a[j:j + block].contiguous())
        out[i:i + block, j:j + block] = blk
        if j != i:
            out[j:j + block, i:i + block] = blk.t()
    if world > 1:
        dist.all_reduce(out, group=group)
    # the diagonal blocks were computed in full: make them exactly symmetric too (upper triangle wins)
    return torch.triu(out) + torch.triu(out, 1).t()


def mmd_cov(m_sr):
    """PointFlow convention on a [n_sample, n_ref] distance matrix.
    MMD = mean over references of the distance to the closest sample; COV = fraction of references that are the
    nearest reference of at least one sample."""
    n_s, n_r = m_sr.shape
    mmd = m_sr.min(dim=0).values.mean()
    cov = torch.unique(m_sr.min(dim=1).indices).numel() / float(n_r)
    return {"mmd": float(mmd), "cov": float(cov)}


def one_nna(m_ss, m_sr, m_rr):
    """Leave-one-out 1-nearest-neighbour accuracy of the two-sample test (0.5 is ideal)."""
    n_s, n_r = m_sr.shape
    top = torch.cat([m_ss, m_sr], dim=1)
    bot = torch.cat([m_sr.t(), m_rr], dim=1)
    full = torch.cat([top, bot], dim=0).clone()
    full.fill_diagonal_(float("inf"))
    label = torch.cat([torch.ones(n_s), torch.zeros(n_r)]).to(full.device)
    pred = label[full.min(dim=1).indices]
    return {"1nna": float((pred == label).float().mean())}


def compute_metrics(samples, refs, metrics=("cd", "emd"), pair_fn=None, group=None, block=256, symmetric=("cd",)):
    """samples [ns,N,3], refs [nr,M,3] replicated on every rank -> {"cd": {...}, "emd": {...}}.
    The S x S and R x R blocks of the 1-NNA test are computed as symmetric matrices (half the pairs) for the metrics
    named in `symmetric` — Chamfer by default; the auction EMD of (a, b) and (b, a) agree only approximately, so it is
    evaluated in full unless asked otherwise."""
    out = {}
    for metric in metrics:
        fn = (pair_fn or {}).get(metric) if isinstance(pair_fn, dict) else pair_fn
        m_sr = pairwise_matrix_sharded(samples, refs, metric, fn, group, block)
        if metric in symmetric:
            m_ss = symmetric_matrix_sharded(samples, metric, fn, group, block)
            m_rr = symmetric_matrix_sharded(refs, metric, fn, group, block)
        else:
            m_ss = pairwise_matrix_sharded(samples, samples, metric, fn, group, block)
            m_rr = pairwise_matrix_sharded(refs, refs, metric, fn, group, block)
        res = mmd_cov(m_sr)
        res.update(one_nna(m_ss, m_sr, m_rr))
        out[metric] = res
    return out
