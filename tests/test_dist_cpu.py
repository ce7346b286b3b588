"""CPU, world_size 2, gloo: host-side sharding / gathering logic of the multi-GPU path (ppvector/parallel.py).
The per-rank compute is the oracle here (CPU); on GPUs the same functions wrap the CUDA kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ppvector.parallel import gather_rows, max_over_ranks, shard_range, sharded_score_rows


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import head as oh
    g = torch.Generator().manual_seed(3)
    trials = torch.randn(11, 192, generator=g)      # ragged: 6 + 5 rows
    enroll = torch.randn(7, 192, generator=g)
    score = lambda A, B: torch.from_numpy(oh.cosine_matrix(A.numpy(), B.numpy())).float()
    full = sharded_score_rows(trials, enroll, score)
    b, e = shard_range(13, rank, world)
    rows = torch.arange(b, e, dtype=torch.float32)[:, None].repeat(1, 3)
    gathered = gather_rows(rows, 13)
    t = max_over_ranks(1.0 + rank, torch.device("cpu"))
    if rank == 0:
        np.savez(out_path, full=full.numpy(), gathered=gathered.numpy(), t=t,
                 ref=oh.cosine_matrix(trials.numpy(), enroll.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    assert np.abs(z["full"] - z["ref"]).max() < 1e-6
    assert np.array_equal(z["gathered"][:, 0], np.arange(13, dtype=np.float32))
    assert float(z["t"]) == 2.0  # max over ranks
