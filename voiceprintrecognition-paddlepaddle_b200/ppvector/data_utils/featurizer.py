"""AudioFeaturizer -- drop-in for ppvector/data_utils/featurizer.py:7-80 of the reference.

Same constructor (``feature_method``, ``method_args``), same ``forward(waveforms, input_lens_ratio=None)``
contract ([B,L] or [L] float32 in [-1,1] -> [B,T,F] float32, per-utterance time mean subtracted, optional
tail mask applied after the mean subtraction), same ``feature_dim`` property.  The arithmetic runs in one
fused CUDA kernel family (``csrc/fbank.cu``) behind ``ppv_fbank_forward``.
"""
import ctypes as C

import torch

from ppvector import _lib


class AudioFeaturizer(torch.nn.Module):
    """reference: featurizer.py:16-31 (constructor dispatch on feature_method)"""

    def __init__(self, feature_method='MelSpectrogram', method_args={}):
        super().__init__()
        self._method_args = dict(method_args or {})
        self._feature_method = feature_method
        self._handle = None
        if feature_method == 'Fbank':
            self._cfg = self._fbank_cfg(self._method_args)
        elif feature_method in ('LogMelSpectrogram', 'MelSpectrogram', 'Spectrogram', 'MFCC'):
            # the paddle.audio.features front ends are "next"; no silent fallback.
            raise _lib.PPVError(f'预处理方法 {feature_method} 尚未在 B200 路径实现 (only Fbank is implemented)')
        else:
            raise Exception(f'预处理方法 {self._feature_method} 不存在!')  # featurizer.py:31

    @staticmethod
    def _fbank_cfg(args):
        lib = _lib.load()
        cfg = _lib.FbankCfg()
        lib.ppv_fbank_default_cfg(C.byref(cfg))
        cfg.n_mels = 23  # paddleaudio kaldi.fbank default when n_mels is not given (cf. feature_dim, featurizer.py:77)
        # paddleaudio.compliance.kaldi.fbank keyword names (featurizer.py:97: Kaldi.fbank(waveform, **kwargs))
        known = {'sr': 'sample_rate', 'n_mels': 'n_mels', 'frame_length': 'frame_length_ms',
                 'frame_shift': 'frame_shift_ms', 'preemphasis_coefficient': 'preemph', 'low_freq': 'low_freq',
                 'high_freq': 'high_freq'}
        for k, v in args.items():
            if k not in known:
                raise _lib.PPVError(f'Fbank argument {k!r} is not supported by the B200 kernel')
            setattr(cfg, known[k], type(getattr(cfg, known[k]))(v))
        return cfg

    def _get_handle(self):
        if self._handle is None:
            lib = _lib.load()
            h = C.c_void_p()
            _lib.check(lib.ppv_fbank_create(C.byref(self._cfg), C.byref(h)), 'ppv_fbank_create')
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().ppv_fbank_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def num_frames(self, num_samples: int) -> int:
        return _lib.load().ppv_fbank_num_frames(self._get_handle(), int(num_samples))

    def forward(self, waveforms, input_lens_ratio=None):
        """reference: featurizer.py:33-60"""
        if not torch.is_tensor(waveforms):
            waveforms = torch.as_tensor(waveforms)
        if waveforms.dim() == 1:
            waveforms = waveforms.unsqueeze(0)
        _lib.require_cuda(waveforms, 'waveforms')
        wav = waveforms.to(torch.float32).contiguous()
        B, L = wav.shape
        lib = _lib.load()
        h = self._get_handle()
        T = lib.ppv_fbank_num_frames(h, L)
        if T <= 0:
            raise _lib.PPVError(f'waveform of {L} samples is shorter than one frame')
        ratio = None
        if input_lens_ratio is not None:
            ratio = torch.as_tensor(input_lens_ratio, dtype=torch.float32, device=wav.device).contiguous()
            assert ratio.numel() == B
        out = torch.empty((B, T, self._cfg.n_mels), dtype=torch.float32, device=wav.device)
        with torch.cuda.device(wav.device):
            _lib.check(lib.ppv_fbank_forward(h, _lib.ptr(wav), _lib.ptr(ratio), B, L, _lib.ptr(out),
                                             _lib.current_stream()), 'ppv_fbank_forward')
        return out

    @property
    def feature_dim(self):
        """reference: featurizer.py:62-80"""
        if self._feature_method == 'LogMelSpectrogram':
            return self._method_args.get('n_mels', 128)
        elif self._feature_method == 'MelSpectrogram':
            return self._method_args.get('n_mels', 64)
        elif self._feature_method == 'Spectrogram':
            return self._method_args.get('n_fft', 512) // 2 + 1
        elif self._feature_method == 'MFCC':
            return self._method_args.get('n_mfcc', 40)
        elif self._feature_method == 'Fbank':
            return self._method_args.get('n_mels', 23)
        else:
            raise Exception('没有{}预处理方法'.format(self._feature_method))
