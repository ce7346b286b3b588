"""GPU: paddle.audio.features front ends (Spectrogram / MelSpectrogram / LogMelSpectrogram / MFCC, SURVEY.md §8 row a2') and
SpecAugment masking (row a14) vs the oracle and the golden fixture.  Tolerances: power/mel 2e-5 of the utterance peak (fp32 FFT);
log-mel / MFCC absolute, stated per test."""
import random

import numpy as np
import pytest
import torch

from oracle import spectral as osp
from ppvector import _lib
from ppvector.data_utils.featurizer import AudioFeaturizer
from ppvector.data_utils.spec_aug import SpecAugmentor

pytestmark = pytest.mark.gpu

CASES = [("Spectrogram", dict(n_fft=512, hop_length=160)), ("MelSpectrogram", dict(sr=16000, n_fft=1024, hop_length=160, n_mels=80)),
         ("LogMelSpectrogram", dict(sr=16000, n_fft=512, hop_length=160, win_length=400, n_mels=80, f_min=20.0)), ("MFCC", dict())]


@pytest.mark.parametrize("method,kw", CASES)
def test_golden(cuda, golden_dir, method, kw):
    g = np.load(f"{golden_dir}/spectral_synth.npz")
    fz = AudioFeaturizer(method, kw)
    out = fz(torch.from_numpy(g["wav"]).to(cuda)).double().cpu().numpy()
    want = g[method]
    assert out.shape == want.shape and out.shape[2] == fz.feature_dim
    if method in ("Spectrogram", "MelSpectrogram"):
        assert np.abs(out - want).max() < 2e-5 * np.abs(want).max()
    elif method == "LogMelSpectrogram":
        assert np.abs(out - want).max() < 2e-3  # dB; the silent tail of utterance 1 sits at the amin floor on both sides
    else:
        assert np.abs(out - want).max() < 2e-3 * 8  # DCT of 64 log-mel values: error grows with sqrt(n_mels)


@pytest.mark.parametrize("method,kw,L", [("MelSpectrogram", dict(sr=16000, n_fft=2048, hop_length=512), 48000),
                                         ("MelSpectrogram", dict(sr=16000, n_fft=256, hop_length=64, n_mels=40, htk=True, norm=None, power=1.0), 8000),
                                         ("Spectrogram", dict(n_fft=1024, hop_length=256, win_length=800, power=2.0), 12345),
                                         ("Spectrogram", dict(n_fft=512, hop_length=128, center=False), 4000),
                                         ("MFCC", dict(sr=16000, n_fft=512, hop_length=160, n_mels=40, n_mfcc=13, f_min=0.0, f_max=7600.0), 16000)])
def test_parameter_sweep_with_tail_mask(cuda, method, kw, L):
    g = torch.Generator().manual_seed(L)
    x = (torch.randn(3, L, generator=g) * 0.1).clamp(-1, 1)
    ratio = torch.tensor([1.0, 0.6, 0.25])
    fz = AudioFeaturizer(method, kw)
    out = fz(x.to(cuda), ratio.to(cuda)).double().cpu()
    want = osp.featurize(x.double(), method, input_lens_ratio=ratio, **kw)
    assert out.shape == want.shape
    tol = 2e-5 * want.abs().max() if method != "MFCC" else 2e-2
    assert (out - want).abs().max() < tol
    T = out.shape[1]
    assert (out[2, int(0.25 * T):] == 0).all() and fz.num_frames(L) == T


def test_unsupported_arguments_fail_loudly():
    with pytest.raises(_lib.PPVError):
        AudioFeaturizer("MelSpectrogram", dict(n_fft=1000))._get_handle()
    with pytest.raises(_lib.PPVError):
        AudioFeaturizer("LogMelSpectrogram", dict(top_db=80.0))
    with pytest.raises(_lib.PPVError):
        AudioFeaturizer("Spectrogram", dict(window="hamming"))


@pytest.mark.parametrize("zero", [True, False])
def test_spec_augment_matches_oracle(cuda, zero):
    g = torch.Generator().manual_seed(5)
    B, T, F = 6, 298, 80
    x = torch.randn(B, T, F, generator=g)
    lens = [298, 298, 200, 150, 298, 17]
    aug = SpecAugmentor(prob=0.7, freq_mask_ratio=0.1, n_freq_masks=2, time_mask_ratio=0.05, n_time_masks=3, inplace=False, replace_with_zero=zero)
    rng = random.Random(1000)
    rows = [aug.draw(lens[b], F, rng) for b in range(B)]
    assert any(r[0] for r in rows) and not all(r[0] for r in rows)
    y = aug.apply(x.to(cuda), rows).cpu()
    for b in range(B):
        r = rows[b]
        if not r[0]:
            assert torch.equal(y[b], x[b])
            continue
        fm = [(r[2 + 2 * i], r[3 + 2 * i]) for i in range(2)]
        tm = [(r[6 + 2 * i], r[7 + 2 * i]) for i in range(3)]
        want = osp.spec_augment_apply(x[b, :lens[b]].numpy(), fm, tm, fill_mean=not zero)
        assert np.abs(y[b, :lens[b]].numpy() - want).max() < 1e-6
        assert torch.equal(y[b, lens[b]:], x[b, lens[b]:])  # padding frames are not touched


def test_spec_augment_config_of_the_reference(cuda):
    # configs/augmentation.yml:36-48: prob 0.5, one frequency mask <= 0.1 * 80 bins, one time mask <= 0.05 * T frames
    aug = SpecAugmentor(prob=0.5, freq_mask_ratio=0.1, n_freq_masks=1, time_mask_ratio=0.05, n_time_masks=1, max_time_warp=0)
    x = torch.randn(64, 298, 80).to(cuda) + 5.0
    y = aug(x.clone(), rng=random.Random(3))
    changed = (y != x).flatten(1).any(1).cpu()
    assert 10 < int(changed.sum()) < 54
    zeros = (y == 0)
    assert int(zeros.all(1).sum(1).max()) <= 8 and int(zeros.all(2).sum(1).max()) <= 14
