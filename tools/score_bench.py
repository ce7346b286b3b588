"""Config 5 of SURVEY.md §8(d): 10^6 cosine scores -- all-pairs [1000 x 1000] (trainer.py:416-423 semantics) and the pair-list form
(idx [10^6, 2] into 10 000 embeddings), embeddings resident in HBM.  Prints one JSON line per form with the roofline the kernel is
bound by (all-pairs: 384 FLOP + 4 B written per pair; pair list: 1536 B read per pair)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "voiceprintrecognition-paddlepaddle_b200"))
from ppvector.metric.cosine import cosine_matrix  # noqa: E402
from ppvector.metric.cosine import cosine_pairlist as cosine_pairs  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1000)
    trial = torch.randn(1000, 192, generator=g).to(dev)
    enroll = torch.randn(1000, 192, generator=g).to(dev)
    ms = timed(lambda: cosine_matrix(trial, enroll), 200)
    ref = torch.nn.functional.normalize(trial.double(), dim=1) @ torch.nn.functional.normalize(enroll.double(), dim=1).T
    err = (cosine_matrix(trial, enroll).double() - ref).abs().max().item()
    print(json.dumps({"form": "all_pairs_1000x1000", "ms": round(ms, 4), "pairs_per_s": round(1e6 / ms * 1e3), "max_abs_err_vs_fp64": err,
                      "GB_per_s_written": round(4e6 / ms / 1e6, 1), "GFLOP_per_s": round(384e6 / ms / 1e6, 1), "bound": "launch latency (0.38 GFLOP, 4 MB)"}))
    table = torch.randn(10000, 192, generator=g).to(dev)
    idx = torch.randint(0, 10000, (1000000, 2), generator=g, dtype=torch.int32).to(dev)
    ms = timed(lambda: cosine_pairs(table, idx), 50)
    got = cosine_pairs(table, idx)[:1000].double()
    a, b = table[idx[:1000, 0].long()].double(), table[idx[:1000, 1].long()].double()
    err = (got - torch.nn.functional.cosine_similarity(a, b)).abs().max().item()
    gbs = 1536e6 / ms / 1e6
    print(json.dumps({"form": "pair_list_1e6_of_10000", "ms": round(ms, 4), "pairs_per_s": round(1e6 / ms * 1e3), "max_abs_err_vs_fp64": err,
                      "algorithmic_GB_per_s": round(gbs, 1), "hbm_peak_GB_per_s": PEAKS.get("hbm_gbs"),
                      "note": "the 7.7 MB table is L2 resident: the gathers are served by L2, not HBM"}))


if __name__ == "__main__":
    main()
