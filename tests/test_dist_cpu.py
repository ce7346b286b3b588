"""CPU, world_size 2, gloo: host-side sharding / gathering logic of the multi-GPU path (ppvector/parallel.py).
The per-rank compute is the oracle here (CPU); on GPUs the same functions wrap the CUDA kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ppvector.parallel import allreduce_flat_grads, gather_rows, max_over_ranks, shard_range, sharded_score_rows, train_sample_indices


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
    # training: one all-reduce over the flat gradient buffer, then the same optimizer step on every rank
    grads = torch.arange(10, dtype=torch.float32) * (rank + 1)
    scale = allreduce_flat_grads(grads)
    idx = train_sample_indices(7, 3, rank, world)
    if rank == 0:
        np.savez(out_path, full=full.numpy(), gathered=gathered.numpy(), t=t,
                 ref=oh.cosine_matrix(trials.numpy(), enroll.numpy()), grads=grads.numpy(), scale=scale, idx=idx)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    assert np.abs(z["full"] - z["ref"]).max() < 1e-6
    assert np.array_equal(z["gathered"][:, 0], np.arange(13, dtype=np.float32))
    assert float(z["t"]) == 2.0  # max over ranks
    assert np.array_equal(z["grads"], np.arange(10, dtype=np.float32) * 3) and float(z["scale"]) == 0.5
    assert len(z["idx"]) == 4


def test_train_sample_indices_cover_the_epoch():
    for n, world in ((7, 2), (10, 3), (64, 8)):
        parts = [train_sample_indices(n, 5, r, world) for r in range(world)]
        assert len({len(p) for p in parts}) == 1  # every rank sees the same number of samples
        assert set(np.concatenate(parts).tolist()) == set(range(n))
        assert not np.array_equal(np.sort(parts[0]), parts[0]) or n < 3  # shuffled
    assert np.array_equal(train_sample_indices(5, 0, 0, 1, shuffle=False), np.arange(5))
    assert allreduce_flat_grads(torch.ones(3)) == 1.0  # no process group: identity
