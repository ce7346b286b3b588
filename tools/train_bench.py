"""Training-step throughput of the ECAPA-TDNN CUDA trainer (SURVEY.md §8d config 3 shape: per-GPU batch 64 x 298 frames, 2796
speakers, AAM margin 0.2, Adam) -- a tuning aid; features resident in HBM.  python tools/train_bench.py [--batch 64] [--frames 298]
[--steps 10] [--once]  (--once: one warm step only, for an ncu launch list).  Under torchrun every rank trains its own batch and the
gradient all-reduce runs over NCCL."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "voiceprintrecognition-paddlepaddle_b200"))
from ppvector.models.ecapa_tdnn import EcapaTdnn  # noqa: E402
from ppvector.train_engine import TrainEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=298)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--speakers", type=int, default=2796)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16"], help="bf16 = train_conf.enable_amp")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    torch.manual_seed(1000)
    eng = TrainEngine(input_size=80, num_speakers=a.speakers, device=dev)
    eng.set_precision(a.precision)
    eng.load_state_dict(EcapaTdnn(input_size=80).state_dict(), torch.nn.init.xavier_uniform_(torch.empty(192, a.speakers)))
    g = torch.Generator().manual_seed(1000 + rank)
    x = torch.randn(a.batch, a.frames, 80, generator=g)
    x = (x - x.mean(1, keepdim=True)).to(dev)
    y = torch.randint(0, a.speakers, (a.batch,), generator=g).to(dev)

    def step():
        loss = eng.forward_backward(x, y, margin=0.2)
        eng.adam_step(lr=1e-3, weight_decay=1e-6, grad_scale=eng.all_reduce_grads())
        return loss

    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    if a.once:
        loss = step()
        torch.cuda.synchronize()
        print(json.dumps({"loss": float(loss)}))
        return
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    t0.record()
    for _ in range(a.steps):
        loss = step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / a.steps
    # breakdown: forward + backward | gradient all-reduce | Adam, CUDA events around each phase of the same step
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(a.steps)]
    if world > 1:
        dist.barrier()
    for i in range(a.steps):
        ev[i][0].record()
        eng.forward_backward(x, y, margin=0.2)
        ev[i][1].record()
        scale = eng.all_reduce_grads()
        ev[i][2].record()
        eng.adam_step(lr=1e-3, weight_decay=1e-6, grad_scale=scale)
        ev[i][3].record()
    torch.cuda.synchronize()
    phases = [sum(e[k].elapsed_time(e[k + 1]) for e in ev) / a.steps for k in range(3)]
    if world > 1:
        tp = torch.tensor(phases, device=dev)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        phases = tp.tolist()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        # algorithmic work: training step ~ 3 x forward, forward 2.857 GFLOP / utterance as executed
        print(json.dumps({"metric": "train_samples_per_s", "value": round(world * a.batch / ms * 1e3, 1), "n_gpus": world, "precision": a.precision, "ms_per_step": round(ms, 3),
                          "batch_per_gpu": a.batch, "frames": a.frames, "speakers": a.speakers, "loss": float(loss),
                          "algorithmic_tflops": round(world * a.batch * 3 * 2.857e9 / (ms * 1e-3) / 1e12, 1),
                          "workspace_GB": round(eng._ws.numel() / 2**30, 2),
                          "breakdown_ms": {"forward_backward": round(phases[0], 3), "grad_all_reduce": round(phases[1], 3), "adam": round(phases[2], 3)},
                          "all_reduce_bytes": int(eng.grads.numel() * 4)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
