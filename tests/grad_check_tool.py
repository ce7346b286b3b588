"""Debug aid (test infrastructure -- it imports oracle/, so it lives under tests/): per-parameter gradient error of the CUDA training step vs the fp64 autograd oracle (prints every tensor)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "voiceprintrecognition-paddlepaddle_b200"))
from oracle import ecapa as oe  # noqa: E402
from oracle import train as ot  # noqa: E402
from ppvector.train_engine import TrainEngine  # noqa: E402

B, T, S = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 40, 37
g = torch.Generator().manual_seed(100 + T)
f = torch.randn(B, T, 80, generator=g, dtype=torch.float64)
f = f - f.mean(1, keepdim=True)
y = torch.randint(0, S, (B,), generator=g)
Wc = (torch.rand(192, S, generator=g, dtype=torch.float64) * 2 - 1) * (6.0 / (192 + S)) ** 0.5
W = oe.make_ecapa_weights(seed=1000, dtype=torch.float64)
taps = {}
loss, grads, stats, logits = ot.train_step_grads(f, y, W, Wc, margin=0.2, taps=taps, layer_taps=True)
eng = TrainEngine(input_size=80, num_speakers=S, device="cuda:0")
eng.load_state_dict(W, Wc)
gl = eng.forward_backward(f.float().cuda(), y.cuda(), margin=0.2)
print("loss", gl.item(), loss.item())


def rd(name, C):
    return eng.read_tap(name, (B, T, C)).double().cpu()


def cmp(label, got, want):
    want = want.transpose(1, 2)
    print(f"{label:30s} rel {((got - want).norm() / want.norm()).item():9.2e}  max|ref| {want.abs().max().item():.3e}")


cmp("d(out3) = D:2", rd("g:D:2", 512), taps["blocks.3"].grad)
cmp("d(out2) = D:1", rd("g:D:1", 512), taps["blocks.2"].grad)
cmp("d(out1) = D:0", rd("g:D:0", 512), taps["blocks.1"].grad)
cmp("d(Y0) = dXt1:0 + D:0", rd("g:dXt1:0", 512) + rd("g:D:0", 512), taps["blocks.0"].grad)
cmp("dOUTCAT[:, 1024:]", rd("g:dOUTCAT", 1536)[..., 1024:], taps["blocks.3"].grad)
for blk in (2, 1, 0):
    p = f"blocks.{blk + 1}"
    cmp(f"{p} dZt2", rd(f"g:dZt2:{blk}", 512), taps[p + ".tdnn2.z"].grad)
    cmp(f"{p} dRC (d res2 out)", rd(f"g:dRC:{blk}", 512), taps[p + ".res2net_block"].grad)
    cmp(f"{p} dZt1", rd(f"g:dZt1:{blk}", 512), taps[p + ".tdnn1.z"].grad)
    cmp(f"{p} dXt1", rd(f"g:dXt1:{blk}", 512), taps["blocks.%d" % blk].grad - taps[p].grad)
o2 = taps["blocks.2"].grad.transpose(1, 2)
print("dXt1:2 + mfa win vs d(out2) - d(out3):", (((rd("g:dXt1:2", 512) + rd("g:dOUTCAT", 1536)[..., 512:1024]) - (o2 - taps["blocks.3"].grad.transpose(1, 2))).norm() / o2.norm()).item())
for name in reversed(list(grads.keys())):
    gw = grads[name]
    gg = eng.view(name, tuple(gw.shape), "grad").double().cpu()
    rel = ((gg - gw).norm() / (gw.norm() + 1e-30)).item()
    cos = ((gg * gw).sum() / (gg.norm() * gw.norm() + 1e-30)).item()
    print(f"{name:55s} rel {rel:9.2e} cos {cos:.6f} |ref| {gw.norm().item():9.3e} |got| {gg.norm().item():9.3e}")
