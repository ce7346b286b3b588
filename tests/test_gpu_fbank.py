"""GPU: fused Fbank kernel (through AudioFeaturizer -> ppv_fbank_forward) vs the oracle and the golden vectors."""
import numpy as np
import pytest
import torch

from oracle import fbank as ofb
from ppvector.data_utils.featurizer import AudioFeaturizer

pytestmark = pytest.mark.gpu

# log-mel domain, fp32 arithmetic on both sides; the intrinsic fp32 noise of Kaldi fbank on real speech is a few
# 1e-4 (see tests/test_oracle_fbank.py), so the kernel is held to the same band against the fp64 oracle.
TOL_MAX = 2e-3
TOL_MEAN = 5e-5


@pytest.fixture(scope="module")
def fz():
    return AudioFeaturizer("Fbank", {"sr": 16000, "n_mels": 80})


@pytest.mark.parametrize("name", ["a_1", "a_2", "b_1", "b_2", "long3s"])
def test_bundled_wavs_vs_golden(cuda, fz, golden_dir, name):
    g = np.load(f"{golden_dir}/fbank_wavs.npz")
    x = g[name + "_pcm"].astype(np.float32) / 32768.0
    ref = g[name + "_fbank"].astype(np.float64)
    ref = ref - ref.mean(0, keepdims=True)  # AudioFeaturizer subtracts the time mean
    out = fz(torch.from_numpy(x).to(cuda)).cpu().numpy()[0].astype(np.float64)
    assert out.shape == ref.shape
    d = np.abs(out - ref)
    assert d.max() < TOL_MAX and d.mean() < TOL_MEAN, (d.max(), d.mean())


def test_batch_mask_and_oracle(cuda, fz, golden_dir):
    g = np.load(f"{golden_dir}/fbank_synth.npz")
    gen = torch.Generator().manual_seed(1000)
    x = (0.1 * torch.randn(4, 48000, generator=gen)).clamp(-1, 1)
    out = fz(x.to(cuda), torch.from_numpy(g["ratio"])).cpu().numpy()
    assert out.shape == (4, 298, 80)
    assert np.abs(out - g["featurizer_masked"]).max() < TOL_MAX
    lens = (g["ratio"] * np.float32(298)).astype(np.int32)
    for b in range(4):
        assert np.all(out[b, lens[b]:] == 0)
    ref64 = ofb.audio_featurizer_fbank(x.numpy(), None, dtype=np.float64, n_mels=80)
    out2 = fz(x.to(cuda)).cpu().numpy()
    d = np.abs(out2 - ref64)
    assert d.max() < TOL_MAX and d.mean() < TOL_MEAN


@pytest.mark.parametrize("L", [400, 559, 560, 4801, 16000, 163840])
def test_ragged_lengths(cuda, fz, L):
    gen = torch.Generator().manual_seed(L)
    x = (0.3 * torch.randn(3, L, generator=gen)).clamp(-1, 1)
    out = fz(x.to(cuda)).cpu().numpy()
    T = ofb.num_frames(L)
    assert out.shape == (3, T, 80)
    ref = ofb.audio_featurizer_fbank(x.numpy(), None, dtype=np.float64, n_mels=80)
    assert np.abs(out - ref).max() < TOL_MAX


def test_silence_hits_log_floor_and_short_input_errors(cuda, fz):
    out = fz(torch.zeros(2, 1600, device=cuda))
    # log(eps) everywhere, minus its own mean: exactly 0 up to the fp32 rounding of an 8-term sum (1 ulp of 15.94)
    assert out.abs().max().item() < 2e-6
    from ppvector._lib import PPVError
    with pytest.raises(PPVError):
        fz(torch.zeros(1, 399, device=cuda))


def test_full_size_properties(cuda, fz):
    """BASELINE config 2 size (256 x 3 s): size-independent properties -- CMN gives zero time-mean per
    (utterance, bin); utterances are independent (batch result == per-utterance result, bit-exact)."""
    gen = torch.Generator().manual_seed(1000)
    x = (0.1 * torch.randn(256, 48000, generator=gen)).clamp(-1, 1).to(cuda)
    out = fz(x)
    assert out.shape == (256, 298, 80) and torch.isfinite(out).all()
    assert out.mean(1).abs().max().item() < 2e-5
    for b in (0, 17, 255):
        assert torch.equal(fz(x[b]), out[b:b + 1])
    # direct oracle comparison of 8 rows of the full-size batch (first / last work items, both ends of the batch)
    rows = [0, 1, 63, 100, 127, 200, 254, 255]
    ref = ofb.audio_featurizer_fbank(x[rows].cpu().numpy(), None, dtype=np.float64, n_mels=80)
    d = np.abs(out[rows].cpu().numpy() - ref)
    assert d.max() < TOL_MAX and d.mean() < TOL_MEAN, (d.max(), d.mean())
