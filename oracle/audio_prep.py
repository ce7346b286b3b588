"""Oracle: waveform preparation / augmentation of the training data path on the CPU (numpy, fp64 statistics).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED for this file: the arithmetic lives in the un-vendored ``yeaudio``
package (requirements.txt:12, call sites ppvector/data_utils/reader.py:8-10, 85-101, 143-163); what is restated here is its RECALLED
behaviour, in the order reader.py applies it:

  reader.py:92-93   resample to the target rate                    (host, ppvector.data_utils.audio.resample)
  reader.py:155-156 SpeedPerturbAugmentor   change_speed(r): new_len = int(len / r);
                                            samples = np.interp(np.linspace(0, len, new_len), np.arange(len), samples)
  reader.py:157-158 VolumePerturbAugmentor  samples *= 10 ** (gain_dB / 20)
  reader.py:159-160 NoisePerturbAugmentor   noise_gain_dB = min(rms_dB(audio) - rms_dB(noise) - snr_dB, 300); audio += noise * 10 ** (g / 20)
                                            (noise tiled from an offset over the utterance)
  reader.py:97-98   normalize(target_db)    gain = min(target_dB - rms_dB(samples), 300) dB, rms_dB = 10 log10(mean(x ** 2))
  reader.py:100-101 crop(duration, mode)    train: random start; eval: start 0
"""
import numpy as np


def rms_db(x):
    ms = float(np.mean(np.asarray(x, dtype=np.float64) ** 2))
    return 10.0 * np.log10(ms) if ms > 0 else -np.inf


def change_speed(x, speed_rate):
    if speed_rate == 1.0:
        return np.asarray(x, dtype=np.float32)
    old = x.shape[0]
    new = int(old / speed_rate)
    return np.interp(np.linspace(0, old, new), np.arange(old), x).astype(np.float32)


def prepare(x, speed_rate=1.0, vol_gain_db=0.0, noise=None, noise_off=0, snr_db=None, target_db=-20.0, normalize=True, crop_start=0,
            crop_len=None, out_len=None):
    """One utterance through speed -> volume -> noise -> dB normalise -> crop -> zero-pad; float64 inside, float32 out."""
    y = change_speed(np.asarray(x, dtype=np.float32), speed_rate).astype(np.float64)
    y = y * 10.0 ** (vol_gain_db / 20.0)
    if noise is not None:
        n = np.asarray(noise, dtype=np.float64)
        seg = n[(noise_off + np.arange(y.shape[0])) % n.shape[0]] if noise_off < n.shape[0] else n[np.arange(y.shape[0]) % n.shape[0]]
        g = min(rms_db(y) - rms_db(seg) - snr_db, 300.0)
        y = y + seg * 10.0 ** (g / 20.0)
    if normalize:
        y = y * 10.0 ** (min(target_db - rms_db(y), 300.0) / 20.0)
    crop_len = y.shape[0] - crop_start if crop_len is None else crop_len
    y = y[crop_start:crop_start + crop_len]
    out_len = y.shape[0] if out_len is None else out_len
    out = np.zeros(out_len, dtype=np.float32)
    out[:y.shape[0]] = y.astype(np.float32)
    return out
