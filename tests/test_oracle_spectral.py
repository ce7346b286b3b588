"""CPU: the STFT front-end oracle against torchaudio's independent implementation of the same (librosa) definitions."""
import numpy as np
import pytest
import torch
import torchaudio

from oracle import spectral as osp


@pytest.mark.parametrize("sr,n_fft,n_mels,f_min,htk", [(22050, 2048, 64, 50.0, False), (16000, 512, 80, 20.0, False), (16000, 1024, 40, 0.0, True)])
def test_mel_matrix_matches_torchaudio(sr, n_fft, n_mels, f_min, htk):
    want = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, f_min, sr / 2.0, n_mels, sr, norm="slaney", mel_scale="htk" if htk else "slaney")
    got = osp.fbank_matrix(sr, n_fft, n_mels, f_min, None, htk, "slaney")
    assert np.abs(got.T - want.double().numpy()).max() < 1e-5 * np.abs(got).max()  # torchaudio evaluates in float32


def test_dct_matches_torchaudio():
    want = torchaudio.functional.create_dct(40, 64, norm="ortho").double().numpy()
    assert np.abs(osp.dct_matrix(40, 64) - want).max() < 5e-6  # torchaudio evaluates in float32


def test_logmel_and_mfcc_match_torchaudio_transforms():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 22050, generator=g, dtype=torch.float64) * 0.1
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=2048, hop_length=512, f_min=50.0, n_mels=64, power=2.0, norm="slaney",
                                               mel_scale="slaney").double()
    want = mel(x)
    got = osp.features(x, "MelSpectrogram")
    assert got.shape == want.shape == (2, 64, 1 + 22050 // 512)
    assert (got - want).abs().max() < 1e-5 * want.abs().max()
    lm = osp.features(x, "LogMelSpectrogram")
    assert torch.allclose(lm, 10 * torch.log10(torch.clamp(want, min=1e-10)), atol=5e-4)  # float32 mel matrix on the torchaudio side
    mf = osp.features(x, "MFCC")
    d = torchaudio.functional.create_dct(40, 64, norm="ortho").double()
    ref = (lm.transpose(1, 2) @ d).transpose(1, 2)
    assert (mf - ref).abs().max() < 1e-5 * ref.abs().max()


def test_golden(golden_dir):
    g = np.load(f"{golden_dir}/spectral_synth.npz")
    x = g["wav"]
    for method, kw in [("Spectrogram", dict(n_fft=512, hop_length=160)), ("MelSpectrogram", dict(sr=16000, n_fft=1024, hop_length=160, n_mels=80)),
                       ("LogMelSpectrogram", dict(sr=16000, n_fft=512, hop_length=160, win_length=400, n_mels=80, f_min=20.0)),
                       ("MFCC", dict())]:
        got = osp.featurize(x, method, **kw).numpy()
        assert np.abs(got - g[method]).max() < 1e-9 * max(1.0, np.abs(g[method]).max())


def test_spec_augment_apply():
    x = np.arange(50, dtype=np.float64).reshape(10, 5)
    y = osp.spec_augment_apply(x, [(1, 2)], [(7, 2)])
    assert (y[:, 1:3] == 0).all() and (y[7:9] == 0).all() and (y[:7, [0, 3, 4]] == x[:7, [0, 3, 4]]).all()
    z = osp.spec_augment_apply(x, [(0, 1)], [], fill_mean=True)
    assert (z[:, 0] == x.mean()).all()
