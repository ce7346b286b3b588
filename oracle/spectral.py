"""Oracle: the STFT front ends of the reference other than Kaldi Fbank, and SpecAugment masking, on the CPU.

TEST INFRASTRUCTURE -- see oracle/__init__.py.

The reference builds ``paddle.audio.features.{Spectrogram, MelSpectrogram, LogMelSpectrogram, MFCC}(**method_args)``
(ppvector/data_utils/featurizer.py:20-27); paddle is not vendored and not installable here, so this restates the
library's published algorithm (paddle/audio/features/layers.py, paddle/audio/functional/functional.py -- the librosa
definitions): ``paddle.signal.stft(center=True, pad_mode='reflect', window=hann(periodic))`` -> ``|X| ** power`` ->
``compute_fbank_matrix`` (slaney or htk mel scale, slaney area norm) -> ``power_to_db`` (10 log10(max(amin, x)) -
10 log10(max(amin, ref)), top_db None) -> ``create_dct(norm='ortho')``.  tests/test_oracle_spectral.py pins the matrices to
torchaudio's independent implementation of the same definitions (melscale_fbanks, create_dct) and the STFT to torch.stft.

``featurize`` adds featurizer.py:43-59 (transpose, time-mean subtraction, tail mask).
``spec_augment_apply`` follows yeaudio's SpecAugmentor as used at reader.py:105-107 (semantics recalled):
frequency masks then time masks, each a [start, start+width) band written with zeros or with the utterance mean.
"""
import math

import numpy as np
import torch

DEFAULTS = dict(sr=22050, n_fft=2048, hop_length=512, win_length=None, power=2.0, center=True, n_mels=64, f_min=50.0, f_max=None,
                htk=False, norm="slaney", ref_value=1.0, amin=1e-10, n_mfcc=40)


def hz_to_mel(f, htk=False):
    f = np.asarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def mel_to_hz(m, htk=False):
    m = np.asarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def fbank_matrix(sr, n_fft, n_mels, f_min, f_max=None, htk=False, norm="slaney"):
    """compute_fbank_matrix -> [n_mels, n_fft // 2 + 1] float64"""
    f_max = sr / 2.0 if f_max is None else f_max
    fftfreqs = np.linspace(0, sr / 2.0, n_fft // 2 + 1)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(f_min, htk), hz_to_mel(f_max, htk), n_mels + 2), htk)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    if norm == "slaney":
        w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w


def dct_matrix(n_mfcc, n_mels):
    """create_dct(norm='ortho') -> [n_mels, n_mfcc] float64"""
    n = np.arange(n_mels, dtype=np.float64)
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    d = np.cos(math.pi / n_mels * (n + 0.5) * k)
    d[0] *= 1.0 / math.sqrt(2.0)
    d *= math.sqrt(2.0 / n_mels)
    return d.T


def stft_power(x, n_fft, hop_length, win_length=None, power=2.0, center=True):
    """x [B,L] float64 tensor -> [B, n_fft//2+1, T]"""
    win_length = win_length or n_fft
    window = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
    X = torch.stft(x, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center, pad_mode="reflect",
                   return_complex=True)
    return X.abs() ** power


def features(x, method, **kw):
    """paddle.audio.features.<method>(**kw)(x): x [B,L] -> [B,F,T] float64"""
    a = dict(DEFAULTS)
    if method == "Spectrogram":
        a.update(n_fft=512, power=1.0)
    a.update(kw)
    x = torch.as_tensor(x, dtype=torch.float64)
    S = stft_power(x, a["n_fft"], a["hop_length"], a["win_length"], a["power"], a["center"])
    if method == "Spectrogram":
        return S
    fb = torch.from_numpy(fbank_matrix(a["sr"], a["n_fft"], a["n_mels"], a["f_min"], a["f_max"], a["htk"], a["norm"]))
    mel = torch.matmul(fb, S)
    if method == "MelSpectrogram":
        return mel
    logmel = 10.0 * torch.log10(torch.clamp(mel, min=a["amin"])) - 10.0 * math.log10(max(a["amin"], a["ref_value"]))
    if method == "LogMelSpectrogram":
        return logmel
    assert method == "MFCC"
    d = torch.from_numpy(dct_matrix(a["n_mfcc"], a["n_mels"]))
    return torch.matmul(logmel.transpose(1, 2), d).transpose(1, 2)


def featurize(x, method, input_lens_ratio=None, **kw):
    """AudioFeaturizer.forward, featurizer.py:43-59: [B,L] -> [B,T,F], time mean removed, tail frames zeroed afterwards"""
    f = features(x, method, **kw).transpose(1, 2)
    f = f - f.mean(1, keepdim=True)
    if input_lens_ratio is not None:
        T = f.shape[1]
        lens = (torch.as_tensor(input_lens_ratio, dtype=torch.float32) * T).to(torch.int32)
        mask = torch.arange(T)[None, :] < lens[:, None]
        f = torch.where(mask[..., None], f, torch.zeros_like(f))
    return f


def spec_augment_apply(x, freq_masks, time_masks, fill_mean=False):
    """x [T,F] numpy; masks = lists of (start, width).  Frequency masks first, then time masks; the fill value is taken once,
    before any mask is written."""
    x = np.array(x, copy=True)
    fill = x.mean() if fill_mean else 0.0
    for f0, w in freq_masks:
        x[:, f0:f0 + w] = fill
    for t0, w in time_masks:
        x[t0:t0 + w, :] = fill
    return x
