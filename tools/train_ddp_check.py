"""Two-rank check of the data-parallel training step (run under torchrun --nproc-per-node 2): after the single all-reduce every
rank holds the SUM of the per-rank gradients, Adam with grad_scale = 1/world keeps the replicas bitwise identical."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "voiceprintrecognition-paddlepaddle_b200"))
from ppvector.models.ecapa_tdnn import EcapaTdnn  # noqa: E402
from ppvector.train_engine import TrainEngine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
S, B, T = 100, 8, 120
torch.manual_seed(1000)
sd = EcapaTdnn(input_size=80).state_dict()
wc = torch.nn.init.xavier_uniform_(torch.empty(192, S))


def batch(r):
    g = torch.Generator().manual_seed(50 + r)
    x = torch.randn(B, T, 80, generator=g)
    return (x - x.mean(1, keepdim=True)).to(dev), torch.randint(0, S, (B,), generator=g).to(dev)


eng = TrainEngine(input_size=80, num_speakers=S, device=dev)
eng.load_state_dict(sd, wc)
x, y = batch(rank)
eng.forward_backward(x, y, margin=0.2)
scale = eng.all_reduce_grads()
reduced = eng.grads.clone()
# reference: the same engine computes every rank's gradient locally and sums them
ref = TrainEngine(input_size=80, num_speakers=S, device=dev)
total = torch.zeros_like(reduced)
for r in range(world):
    ref.load_state_dict(sd, wc)
    xr, yr = batch(r)
    ref.forward_backward(xr, yr, margin=0.2)
    total += ref.grads
err = ((reduced - total).norm() / total.norm()).item()
eng.adam_step(lr=1e-3, weight_decay=1e-6, grad_scale=scale)
gathered = [torch.empty_like(eng.params) for _ in range(world)]
dist.all_gather(gathered, eng.params)
same = all(torch.equal(gathered[0], g) for g in gathered)
if rank == 0:
    print(json.dumps({"world": world, "grad_scale": scale, "allreduce_rel_err": err, "replicas_identical_after_adam": same}))
assert err < 1e-6 and same and scale == 1.0 / world
dist.destroy_process_group()
