"""CPU tests (gloo, world_size 2) of the host-side logic of the block-sharded all-pairs evaluation
(pcfm/evaluation.py): row partition, all-gather of a reference set loaded sharded, gather of the row blocks, and the
MMD / COV / 1-NNA reductions.  The per-pair kernel is injected (stock-PyTorch CPU statement from oracle/torch_cpu.py):
the CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcfm import evaluation as ev
from oracle import torch_cpu as tc


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [ev.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


def test_metrics_on_known_matrices():
    m_sr = torch.tensor([[0.1, 0.9, 0.8], [0.7, 0.2, 0.9]])
    r = ev.mmd_cov(m_sr)
    assert abs(r["mmd"] - (0.1 + 0.2 + 0.8) / 3) < 1e-6 and abs(r["cov"] - 2 / 3) < 1e-6
    # two well separated clusters: 1-NN always finds the own set => accuracy 1.0; identical sets => 0.0 on ties
    a = torch.zeros(3, 3) + 0.01
    far = torch.ones(3, 3) * 5
    assert ev.one_nna(a, far, a)["1nna"] == 1.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        S = torch.randn(7, 64, 3, generator=g) * 0.3      # 7 rows over 2 ranks: uneven split (4 + 3)
        R = torch.randn(5, 48, 3, generator=g) * 0.3
        # reference set arrives sharded -> all-gather
        lo, hi = ev.shard_range(R.shape[0], rank, world)
        R_full = ev.all_gather_clouds(R[lo:hi].contiguous())
        assert torch.equal(R_full, R)
        full = ev.pairwise_matrix_sharded(S, R_full, "cd", pair_fn=tc.chamfer_allpairs, block=2)
        lo2, hi2, local = ev.pairwise_matrix_sharded(S, R_full, "cd", pair_fn=tc.chamfer_allpairs, block=3, gather=False)
        res = ev.compute_metrics(S, R_full, metrics=("cd",), pair_fn=tc.chamfer_allpairs, block=4)
        res_full = ev.compute_metrics(S, R_full, metrics=("cd",), pair_fn=tc.chamfer_allpairs, block=4, symmetric=())
        sym = ev.symmetric_matrix_sharded(S, "cd", pair_fn=tc.chamfer_allpairs, block=3)   # 3 + 3 + 1 rows: 6 blocks
        torch.save({"full": full, "lo": lo2, "hi": hi2, "local": local, "res": res, "res_full": res_full, "sym": sym},
                   os.path.join(out_dir, f"r{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_allpairs_world2_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    S = torch.randn(7, 64, 3, generator=g) * 0.3
    R = torch.randn(5, 48, 3, generator=g) * 0.3
    want = tc.chamfer_allpairs(S, R)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for r, o in enumerate(outs):
        assert torch.allclose(o["full"], want, rtol=1e-6, atol=1e-7), f"rank {r} gathered matrix"
        assert (o["lo"], o["hi"]) == ev.shard_range(7, r, world)
        assert torch.allclose(o["local"], want[o["lo"]:o["hi"]], rtol=1e-6, atol=1e-7)
    assert outs[0]["res"] == outs[1]["res"]
    # S x S computed as a symmetric matrix (upper-triangle blocks dealt to the ranks, one all-reduce, mirrored)
    want_ss = tc.chamfer_allpairs(S, S)
    for r, o in enumerate(outs):
        assert torch.equal(o["sym"], o["sym"].t()), f"rank {r}: not symmetric"
        assert torch.allclose(o["sym"], want_ss, rtol=1e-6, atol=1e-7), f"rank {r} symmetric matrix"
        for k, v in o["res_full"]["cd"].items():   # half the pairs, same metrics
            assert abs(v - o["res"]["cd"][k]) < 1e-6
    single = ev.compute_metrics(S, R, metrics=("cd",), pair_fn=tc.chamfer_allpairs, block=4)
    for k, v in single["cd"].items():
        assert abs(v - outs[0]["res"]["cd"][k]) < 1e-6


def test_torch_cpu_chamfer_allpairs_matches_c_oracle():
    from oracle import cpu_ref
    g = torch.Generator().manual_seed(1)
    S = torch.randn(3, 40, 3, generator=g); R = torch.randn(2, 50, 3, generator=g)
    got = tc.chamfer_allpairs(S, R).numpy()
    for i in range(3):
        for j in range(2):
            d1, d2, _, _ = cpu_ref.chamfer_forward(S[i:i + 1].numpy(), R[j:j + 1].numpy())
            assert abs(got[i, j] - (d1.mean() + d2.mean())) < 1e-5
