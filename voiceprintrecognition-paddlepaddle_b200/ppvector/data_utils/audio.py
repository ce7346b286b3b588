"""Minimal audio preparation in front of the hot path (the reference delegates this to the un-vendored
``yeaudio.audio.AudioSegment``: ppvector/predict.py:189-216, ppvector/data_utils/reader.py:85-104).

Only what the drop-in entry points need: PCM16/float WAV decode with the standard library, float32 samples in
[-1,1), dB normalisation to ``target_db`` (gain = target - 10 log10(mean(x^2)); x *= 10^(gain/20)), resampling
by polyphase filtering.  GPU-side audio prep is SURVEY.md §8(f) rank 2 ("next")."""
import io
import wave

import numpy as np


def read_wav(source):
    """path | bytes | file object -> (float32 mono samples in [-1,1), sample_rate)"""
    if isinstance(source, (bytes, bytearray)):
        source = io.BytesIO(source)
    with wave.open(source, 'rb') as w:
        n, sw, ch, sr = w.getnframes(), w.getsampwidth(), w.getnchannels(), w.getframerate()
        raw = w.readframes(n)
    if sw == 2:
        x = np.frombuffer(raw, dtype='<i2').astype(np.float32) / 32768.0
    elif sw == 4:
        x = np.frombuffer(raw, dtype='<i4').astype(np.float32) / 2147483648.0
    elif sw == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f'unsupported sample width {sw}')
    if ch > 1:
        x = x.reshape(-1, ch).mean(axis=1)
    return np.ascontiguousarray(x, dtype=np.float32), sr


def resample(x, sr_from, sr_to):
    if sr_from == sr_to:
        return x
    from math import gcd
    from scipy.signal import resample_poly
    g = gcd(int(sr_from), int(sr_to))
    return resample_poly(x, int(sr_to) // g, int(sr_from) // g).astype(np.float32)


def normalize_db(x, target_db=-20.0, max_gain_db=300.0):
    mean_square = float(np.mean(x.astype(np.float64) ** 2))
    rms_db = 10.0 * np.log10(mean_square) if mean_square > 0.0 else -np.inf  # digital silence cannot be normalised (gain = +inf)
    gain = target_db - rms_db
    if gain > max_gain_db:
        raise ValueError(f'无法将音频归一化到 {target_db} dB: 增益 {gain} dB 超过 max_gain_db')
    return (x * np.float32(10.0 ** (gain / 20.0))).astype(np.float32)


class AudioSegment:
    """The slice of yeaudio.audio.AudioSegment the predictor uses: samples / sample_rate / duration /
    resample / normalize / from_file / from_ndarray / from_bytes."""

    def __init__(self, samples, sample_rate):
        self.samples = np.ascontiguousarray(samples, dtype=np.float32)
        self.sample_rate = int(sample_rate)

    @classmethod
    def from_file(cls, f):
        return cls(*read_wav(f))

    @classmethod
    def from_bytes(cls, b):
        return cls(*read_wav(b))

    @classmethod
    def from_ndarray(cls, data, sample_rate=16000):
        data = np.asarray(data)
        if data.dtype == np.int16:
            data = data.astype(np.float32) / 32768.0
        return cls(data.astype(np.float32), sample_rate)

    @property
    def duration(self):
        return self.samples.shape[0] / float(self.sample_rate)

    def resample(self, target_sample_rate):
        self.samples = resample(self.samples, self.sample_rate, target_sample_rate)
        self.sample_rate = int(target_sample_rate)

    def normalize(self, target_db=-20.0, max_gain_db=300.0):
        self.samples = normalize_db(self.samples, target_db, max_gain_db)
