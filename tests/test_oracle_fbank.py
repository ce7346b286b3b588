"""CPU: the Fbank oracle against the golden vectors (torchaudio kaldi.fbank outputs on the reference's
bundled wavs and on seeded synthetic audio) -- this is what pins the oracle."""
import numpy as np
import pytest

from oracle import fbank

# fp32 Kaldi fbank has an intrinsic noise floor of a few 1e-4 in the log domain (DC removal and
# pre-emphasis cancel most of the energy in the high bins of low-passed speech); fp64-vs-torchaudio(fp32)
# differences measured at mint time were <= 6.4e-4.
TOL_MAX = 2e-3
TOL_MEAN = 5e-5


@pytest.mark.parametrize("name,frames", [("a_1", 365), ("a_2", 218), ("b_1", 502), ("b_2", 516), ("long3s", 298)])
def test_bundled_wavs(golden_dir, name, frames):
    g = np.load(f"{golden_dir}/fbank_wavs.npz")
    x = g[name + "_pcm"].astype(np.float32) / 32768.0
    ref = g[name + "_fbank"]
    assert ref.shape == (frames, 80)  # expected snip_edges frame counts
    assert fbank.num_frames(len(x)) == frames
    for dt in (np.float32, np.float64):
        out = fbank.kaldi_fbank(x, n_mels=80, dtype=dt)
        d = np.abs(out.astype(np.float64) - ref)
        assert d.max() < TOL_MAX and d.mean() < TOL_MEAN, (name, dt, d.max(), d.mean())


def test_synthetic_batch_and_mask(golden_dir):
    import torch
    g = np.load(f"{golden_dir}/fbank_synth.npz")
    gen = torch.Generator().manual_seed(1000)
    x = (0.1 * torch.randn(4, 48000, generator=gen)).clamp(-1, 1).numpy()
    raw = np.stack([fbank.kaldi_fbank(u, n_mels=80) for u in x])
    assert np.abs(raw - g["fbank"]).max() < TOL_MAX
    feat = fbank.audio_featurizer_fbank(x, g["ratio"], n_mels=80)
    assert feat.shape == (4, 298, 80)
    assert np.abs(feat - g["featurizer_masked"]).max() < TOL_MAX
    # mask semantics (featurizer.py:48-59): frames >= int(ratio*T) are exactly zero, mean was taken over all T
    lens = (g["ratio"] * np.float32(298)).astype(np.int32)
    for b in range(4):
        assert np.all(feat[b, lens[b]:] == 0)
    unmasked = fbank.audio_featurizer_fbank(x, None, n_mels=80)
    assert np.abs(unmasked.mean(axis=1)).max() < 1e-4


def test_edge_cases():
    assert fbank.num_frames(399) == 0 and fbank.num_frames(400) == 1 and fbank.num_frames(559) == 1
    assert fbank.num_frames(560) == 2 and fbank.num_frames(48000) == 298
    assert fbank.kaldi_fbank(np.zeros(100, np.float32)).shape == (0, 80)
    # silence hits the log floor
    z = fbank.kaldi_fbank(np.zeros(800, np.float32))
    assert np.allclose(z, np.log(fbank.FLT_EPS))
    w = fbank.mel_banks(80, 512, 16000.0)
    assert w.shape == (80, 257) and (w > 0).sum() == 501 and np.all(w[:, 0] == 0) and np.all(w[:, 256] == 0)


def test_db_normalize():
    rng = np.random.default_rng(0)
    x = (0.05 * rng.standard_normal(16000)).astype(np.float32)
    y = fbank.db_normalize(x, -20.0)
    assert abs(10 * np.log10(np.mean(y.astype(np.float64) ** 2)) + 20.0) < 1e-3
