"""GPU: batched waveform preparation / augmentation (csrc/audio_prep.cu, ppvector.data_utils.audio_batch) against oracle/audio_prep.py
(recalled yeaudio semantics -- see that file's header), the ragged Fbank against the per-utterance featuriser, and the batched dataset
path against the per-item path."""
import random

import numpy as np
import pytest
import torch

from oracle import audio_prep as oap
from ppvector.data_utils.audio_batch import WaveAugmentor, prepare_batch
from ppvector.data_utils.featurizer import AudioFeaturizer

pytestmark = pytest.mark.gpu


def waves(seed, lens):
    rng = np.random.default_rng(seed)
    return [(0.1 * rng.standard_normal(n) * (1 + 0.5 * np.sin(np.arange(n) / 700.0))).astype(np.float32) for n in lens]


def test_prepare_batch_matches_oracle(cuda):
    ws = waves(1, [48000, 30011, 70000, 16000, 52345])
    noise = (0.05 * np.random.default_rng(2).standard_normal(20000)).astype(np.float32)
    bank = torch.from_numpy(noise).to(cuda)
    draws = [dict(speed_rate=1.0, vol_gain_db=0.0, noise=None, snr_db=0.0),
             dict(speed_rate=0.9, vol_gain_db=-7.5, noise=None, snr_db=0.0),
             dict(speed_rate=1.1, vol_gain_db=4.0, noise=(1234, 20000 - 1234), snr_db=15.0),
             dict(speed_rate=1.0, vol_gain_db=0.0, noise=(0, 20000), snr_db=30.0),
             None]
    crops = [(0, None), (100, 20000), (5000, 48000), (0, None), (4345, 48000)]
    out, lens = prepare_batch(ws, draws, crops, target_db=-20.0, normalize=True, noise_bank=bank, device=cuda)
    assert out.shape == (5, max(lens))
    got = out.cpu().numpy()
    for b in range(5):
        d = draws[b] or {}
        nz = None
        if d.get("noise") is not None:
            off, n = d["noise"]
            nz = noise[off:off + n]
        want = oap.prepare(ws[b], d.get("speed_rate", 1.0), d.get("vol_gain_db", 0.0), nz, 0, d.get("snr_db"), -20.0, True, crops[b][0],
                           crops[b][1], out_len=max(lens))
        assert lens[b] == (crops[b][1] if crops[b][1] is not None else len(oap.change_speed(ws[b], d.get("speed_rate", 1.0))) - crops[b][0])
        err = np.abs(got[b] - want).max()
        assert err < 2e-6, (b, err)
        assert np.all(got[b, lens[b]:] == 0)
    # the -20 dB target is met on the un-cropped rows
    assert abs(oap.rms_db(got[0]) + 20.0) < 1e-3 and abs(oap.rms_db(got[3][:16000]) + 20.0) < 1e-3
    # no normalisation: volume gain only
    out2, _ = prepare_batch(ws[:1], [dict(speed_rate=1.0, vol_gain_db=6.0, noise=None, snr_db=0)], [(0, None)], normalize=False, device=cuda)
    assert np.abs(out2.cpu().numpy()[0] - ws[0] * np.float32(10 ** (6.0 / 20))).max() < 1e-6


def test_augmentor_draw_order_and_ranges(cuda, tmp_path):
    import wave
    nd = tmp_path / "noise"
    nd.mkdir()
    for i, n in enumerate([8000, 64000]):
        x = (np.random.default_rng(i).standard_normal(n) * 3000).astype("<i2")
        with wave.open(str(nd / f"n{i}.wav"), "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(16000)
            w.writeframes(x.tobytes())
    conf = {"speed": {"prob": 1.0, "speed_perturb_3_class": True}, "volume": {"prob": 1.0, "min_gain_dBFS": -15, "max_gain_dBFS": 15},
            "noise": {"prob": 1.0, "noise_dir": str(nd), "min_snr_dB": 10, "max_snr_dB": 50}, "reverb": {"prob": 0.5, "reverb_dir": "nowhere"}}
    aug = WaveAugmentor(conf, num_speakers=100, device=cuda)
    assert aug.noise_bank.numel() == 72000 and aug.noise_clips == [(0, 8000), (8000, 64000)]
    rng = random.Random(3)
    seen = set()
    for _ in range(50):
        d = aug.draw(48000, 7, rng)
        k = (1.0, 0.9, 1.1).index(d["speed_rate"])
        assert d["spk_id"] == 7 + 100 * k and -15 <= d["vol_gain_db"] <= 15 and 10 <= d["snr_db"] <= 50
        off, n = d["noise"]
        assert 0 <= off and off + n <= 72000
        seen.add(k)
    assert seen == {0, 1, 2}
    assert WaveAugmentor({"speed": {"prob": 0.0}}, device=cuda).draw(100, 1, rng) == dict(speed_rate=1.0, spk_id=1, vol_gain_db=0.0, noise=None, snr_db=0.0)


def test_ragged_fbank_equals_per_utterance(cuda):
    fz = AudioFeaturizer("Fbank", {"sr": 16000, "n_mels": 80})
    lens = [48000, 16000, 400, 30011, 47999]
    ws = waves(5, lens)
    batch = torch.zeros(len(lens), max(lens))
    for b, w in enumerate(ws):
        batch[b, :len(w)] = torch.from_numpy(w)
    feats, frames = fz.forward_ragged(batch.to(cuda), lens)
    assert frames == [fz.num_frames(n) for n in lens] and feats.shape == (5, max(frames), 80)
    for b, w in enumerate(ws):
        alone = fz(torch.from_numpy(w).to(cuda))[0]
        assert torch.equal(feats[b, :frames[b]], alone), b  # same kernels, same per-utterance arithmetic: bit-exact
        assert (feats[b, frames[b]:] == 0).all()


def test_dataset_batch_path_equals_per_item_path(cuda, tmp_path):
    import wave

    from ppvector.data_utils.collate_fn import collate_fn
    from ppvector.data_utils.reader import PPVectorDataset
    lens = [20000, 70000, 48000, 9000]
    lines = []
    for i, w in enumerate(waves(9, lens)):
        p = tmp_path / f"u{i}.wav"
        with wave.open(str(p), "wb") as f:
            f.setnchannels(1)
            f.setsampwidth(2)
            f.setframerate(16000)
            f.writeframes((w * 32767).astype("<i2").tobytes())
        lines.append(f"{p}\t{i}\n")
    lst = tmp_path / "list.txt"
    lst.write_text("".join(lines))
    fz = AudioFeaturizer("Fbank", {"sr": 16000, "n_mels": 80})
    ds = PPVectorDataset(str(lst), fz, mode="eval", device=cuda)  # eval: crop from 0, no augmentation -> deterministic
    f1, l1, n1 = ds.load_batch(range(4))
    f2, l2, n2 = collate_fn([ds[i] for i in range(4)])
    assert torch.equal(l1, l2) and torch.equal(n1, n2)
    assert torch.allclose(f1, f2, atol=2e-5), (f1 - f2).abs().max()
    assert f1.shape[1] == 298  # the 70000-sample utterance is cropped to max_duration 3 s
