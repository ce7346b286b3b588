"""Data-parallel helpers for embedding extraction / scoring across the GPUs of one node.

The hot path shards by utterance with NO data-path collective: every rank embeds its own slice
with its own stream; the only communication is an optional gather of the [N,192] embeddings at the end (and the
replication of the enrolment matrix for scoring).  One process per GPU, ``torch.distributed`` (NCCL on GPUs;
the same code runs on gloo for the CPU tests of the host logic)."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of n items for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather row shards produced with shard_range back into the [n_total, D] matrix (every rank gets it)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    maxn = max(e - b for b, e in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: e - b] for o, (b, e) in zip(outs, sizes)], dim=0)


def max_over_ranks(seconds: float, device) -> float:
    """Timing rule of bench.py: a multi-GPU number is the MAX over ranks."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_score_rows(trial_emb: torch.Tensor, enroll_emb: torch.Tensor, score_fn) -> torch.Tensor:
    """Config 5 layout: trial rows sharded over ranks, enrolment matrix replicated; returns the full [M,N] score
    matrix on every rank.  `score_fn(A, B) -> [len(A), len(B)]` is ppvector.metric.cosine.cosine_matrix on GPUs."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    b, e = shard_range(trial_emb.shape[0], rank, world)
    local = score_fn(trial_emb[b:e], enroll_emb) if e > b else trial_emb.new_zeros((0, enroll_emb.shape[0]))
    return gather_rows(local, trial_emb.shape[0])


def allreduce_flat_grads(grads: torch.Tensor) -> float:
    """Training: ONE all-reduce(sum) over the flat gradient buffer, in place; returns the factor the optimizer must
    apply to the summed gradient (1 / world size) -- the reference's fleet.distributed_model averaging (trainer.py:318-320)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return 1.0
    dist.all_reduce(grads, op=dist.ReduceOp.SUM)
    return 1.0 / dist.get_world_size()


def train_sample_indices(n: int, epoch: int, rank: int, world: int, shuffle: bool = True):
    """paddle.io.DistributedBatchSampler's index stream: one permutation per epoch (seeded by the epoch, identical on every rank),
    padded by wrapping to a multiple of the world size, rank r takes every world-th index."""
    import numpy as np
    idx = np.arange(n)
    if shuffle:
        np.random.RandomState(epoch).shuffle(idx)
    total = -(-n // world) * world
    return np.concatenate([idx, idx[: total - n]])[rank::world]
