"""Oracle: Kaldi log-Mel filterbank + the AudioFeaturizer wrapper (CPU, numpy).

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows
  * ppvector/data_utils/featurizer.py:88-101  (KaldiFbank.forward: per-utterance
    ``paddleaudio.compliance.kaldi.fbank(waveform[1,L], sr=16000, n_mels=80)``; every
    other argument is paddleaudio's default: frame_length 25 ms, frame_shift 10 ms,
    dither 0, preemphasis 0.97, remove_dc_offset, povey window, round_to_power_of_two,
    snip_edges, low_freq 20, high_freq 0 (= Nyquist), use_power, use_log_fbank,
    use_energy False, no VTLN),
  * ppvector/data_utils/featurizer.py:33-60   (AudioFeaturizer.forward: [B,F,T]->[B,T,F],
    subtract the per-utterance time mean, optional tail mask).

paddleaudio is not vendored in /root/reference and not installable here; its
kaldi module is a port of ``torchaudio/compliance/kaldi.py`` (present in this image),
whose algorithm is restated below and against which this file is pinned
(tests/test_oracle_fbank.py).  Known possible divergence: the log floor.  torchaudio
uses ``finfo(float32).eps`` = 1.1920929e-07; paddleaudio is believed to use the same
constant via ``paddle.finfo``-less literal 1.1920928955078125e-07 -- kept as the
``log_floor`` parameter so a Paddle box can settle it.
"""
import math

import numpy as np

FLT_EPS = float(np.finfo(np.float32).eps)  # 1.1920929e-07


def next_pow2(x: int) -> int:
    return 1 if x == 0 else 2 ** (x - 1).bit_length()


def num_frames(num_samples: int, window_size: int = 400, window_shift: int = 160) -> int:
    """snip_edges=True frame count (torchaudio kaldi.py:_get_strided)."""
    if num_samples < window_size:
        return 0
    return 1 + (num_samples - window_size) // window_shift


def povey_window(window_size: int, dtype=np.float64) -> np.ndarray:
    """hann(periodic=False) ** 0.85 (torchaudio kaldi.py:_feature_window_function)."""
    n = np.arange(window_size, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * math.pi * n / (window_size - 1))
    return (hann ** 0.85).astype(dtype)


def mel_scale(f):
    return 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_banks(num_bins: int, padded_window: int, sample_freq: float,
              low_freq: float = 20.0, high_freq: float = 0.0, dtype=np.float32) -> np.ndarray:
    """Triangular mel filters, [num_bins, padded_window//2 + 1] (last column zero).

    torchaudio kaldi.py:get_mel_banks with vtln_warp_factor == 1.  torchaudio evaluates
    this in float32; we evaluate in float32 as well so the weights agree to the ulp.
    """
    num_fft_bins = padded_window // 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    f32 = np.float32
    fft_bin_width = sample_freq / padded_window
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = np.arange(num_bins, dtype=f32)[:, None]
    left = (f32(mel_low) + b * f32(delta)).astype(f32)
    center = (f32(mel_low) + (b + f32(1.0)) * f32(delta)).astype(f32)
    right = (f32(mel_low) + (b + f32(2.0)) * f32(delta)).astype(f32)
    freqs = (f32(fft_bin_width) * np.arange(num_fft_bins, dtype=f32)).astype(f32)
    mel = (f32(1127.0) * np.log(f32(1.0) + freqs / f32(700.0))).astype(f32)[None, :]
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = np.maximum(f32(0.0), np.minimum(up, down)).astype(dtype)
    return np.pad(bins, ((0, 0), (0, 1)))


def kaldi_fbank(waveform: np.ndarray, sr: int = 16000, n_mels: int = 80,
                frame_length_ms: float = 25.0, frame_shift_ms: float = 10.0,
                preemph: float = 0.97, low_freq: float = 20.0, high_freq: float = 0.0,
                log_floor: float = FLT_EPS, dtype=np.float32) -> np.ndarray:
    """One utterance: waveform [L] -> log-mel [T, n_mels].

    Mirrors torchaudio kaldi.py:fbank (:591-647) / _get_window (:154-217) with the
    defaults the reference uses (featurizer.py:97).  ``dtype`` selects the arithmetic
    precision (float32 = what the reference computes in; float64 = "truth").
    """
    x = np.asarray(waveform, dtype=dtype).reshape(-1)
    win = int(sr * frame_length_ms * 0.001)
    shift = int(sr * frame_shift_ms * 0.001)
    padded = next_pow2(win)
    T = num_frames(x.shape[0], win, shift)
    if T == 0:
        return np.zeros((0, n_mels), dtype=dtype)
    idx = np.arange(T)[:, None] * shift + np.arange(win)[None, :]
    frames = x[idx]                                              # as_strided framing
    frames = frames - frames.mean(axis=1, keepdims=True)         # remove_dc_offset
    prev = np.concatenate([frames[:, :1], frames[:, :-1]], axis=1)  # replicate-pad left
    frames = frames - dtype(preemph) * prev                      # pre-emphasis
    frames = frames * povey_window(win, dtype)[None, :]
    frames = np.pad(frames, ((0, 0), (0, padded - win)))         # zero-pad to 512
    spec = np.fft.rfft(frames.astype(np.float64 if dtype == np.float64 else np.float32), axis=1)
    power = (spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2).astype(dtype)
    banks = mel_banks(n_mels, padded, float(sr), low_freq, high_freq, dtype=np.float32).astype(dtype)
    mel = power @ banks.T
    return np.log(np.maximum(mel, dtype(log_floor))).astype(dtype)


def audio_featurizer_fbank(waveforms: np.ndarray, input_lens_ratio=None, dtype=np.float32,
                           **fbank_args) -> np.ndarray:
    """AudioFeaturizer('Fbank').forward  (featurizer.py:33-60): [B,L] -> [B,T,F].

    CMN takes the mean over ALL T frames, including frames that came from zero padding
    (quirk kept, SURVEY.md quirks register); the tail mask (frames t >= int(ratio*T) -> 0)
    is applied AFTER the mean subtraction (featurizer.py:48-59).
    """
    w = np.asarray(waveforms)
    if w.ndim == 1:
        w = w[None, :]
    feats = np.stack([kaldi_fbank(u, dtype=dtype, **fbank_args) for u in w])   # [B,T,F]
    feats = feats - feats.mean(axis=1, keepdims=True)
    if input_lens_ratio is not None:
        T = feats.shape[1]
        # paddle: (ratio * T).astype(int32) -- float32 multiply then truncate
        lens = (np.asarray(input_lens_ratio, dtype=np.float32) * np.float32(T)).astype(np.int32)
        mask = np.arange(T)[None, :] < lens[:, None]
        feats = np.where(mask[:, :, None], feats, 0).astype(dtype)
    return feats.astype(dtype)


def db_normalize(samples: np.ndarray, target_db: float = -20.0, max_gain_db: float = 300.0) -> np.ndarray:
    """yeaudio AudioSegment.normalize (called at reader.py:97-98, predict.py:214-215).

    yeaudio is not vendored; recalled semantics: rms_db = 10*log10(mean(x^2)),
    gain = target_db - rms_db (error if > max_gain_db), x *= 10^(gain/20).
    """
    x = np.asarray(samples, dtype=np.float32)
    mean_square = np.mean(x.astype(np.float64) ** 2)
    rms_db = 10.0 * math.log10(max(mean_square, 1e-30))
    gain = target_db - rms_db
    if gain > max_gain_db:
        raise ValueError("gain exceeds max_gain_db")
    return (x * np.float32(10.0 ** (gain / 20.0))).astype(np.float32)
