"""Throughput of the waveform -> embedding path with 1 / 2 / 3 batches in flight (PPVectorPredictor.embed_resident_stream and
extract_embeddings_stream lanes): how much of the SM idle time at kernel tails / between dependent launches a second compute lane recovers.
python tools/lanes_bench.py [--steps 100]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "voiceprintrecognition-paddlepaddle_b200"))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    import yaml
    from ppvector.predict import PPVectorPredictor
    dev = torch.device("cuda", 0)
    cfg = yaml.load(open(os.path.join(ROOT, "configs", "ecapa_tdnn.yml")), Loader=yaml.FullLoader)
    Wts = {k: v.numpy() for k, v in bench.seeded_ecapa_weights().items()}
    pred = PPVectorPredictor(cfg, model_path=None, use_gpu=True, state_dict=Wts)
    wavs = [bench.synth_wave(bench.BATCH, 1000 + i).to(dev) for i in range(2)]
    host = [bench.synth_wave(bench.BATCH, 2000 + i).pin_memory() for i in range(2)]
    ref = [e.clone() for e in pred.embed_resident_stream([wavs[0], wavs[1]], lanes=1)]
    for lanes in (1, 2, 3):
        same = all(torch.equal(o, ref[i % 2]) for i, o in enumerate(pred.embed_resident_stream([wavs[i % 2] for i in range(6)], lanes=lanes)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in pred.embed_resident_stream((wavs[i % 2] for i in range(a.steps)), lanes=lanes):
            pass
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        for _ in pred.extract_embeddings_stream((host[i % 2] for i in range(4)), lanes=lanes):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for out in pred.extract_embeddings_stream((host[i % 2] for i in range(a.steps)), lanes=lanes):
            pass
        torch.cuda.synchronize()
        e2e = (time.perf_counter() - t0) / a.steps
        print(json.dumps({"lanes": lanes, "resident_ms_per_step": round(ms, 4), "resident_utt_per_s": round(bench.BATCH / ms * 1e3),
                          "e2e_ms_per_step": round(e2e * 1e3, 4), "e2e_utt_per_s": round(bench.BATCH / e2e), "bitwise_equal_to_one_lane": same}), flush=True)


if __name__ == "__main__":
    main()
