"""PPVectorPredictor -- drop-in for the embedding / contrast surface of ppvector/predict.py:24-283.

Same constructor arguments and the same ``predict`` / ``predict_batch`` / ``contrast`` contracts
(numpy in, numpy out).  Differences that are the point of this build:
  * the waveform batch goes host -> device once (pinned staging buffer) and waveform -> Fbank -> ECAPA-TDNN ->
    embedding is ONE library call (``ppv_model_forward_wav``); the reference featurises per utterance in a Python
    loop and runs the model in chunks of 32 (predict.py:262-267);
  * cosine scoring runs on the GPU (``ppvector.metric.cosine``);
  * ``use_gpu=False`` raises: the B200 build has no CPU path.
Not carried over (out of scope): the enrolment DB, recognition and diarization.
"""
import os
from io import BufferedReader

import numpy as np
import torch
import yaml
from loguru import logger

from ppvector import _lib
from ppvector.data_utils.audio import AudioSegment
from ppvector.data_utils.featurizer import AudioFeaturizer
from ppvector.metric.cosine import cosine_matrix
from ppvector.models import build_model
from ppvector.utils.checkpoint import load_state_dict_file
from ppvector.utils.utils import dict_to_object, print_arguments


class PPVectorPredictor:
    def __init__(self, configs, threshold=0.6, audio_db_path=None, model_path='models/EcapaTdnn_Fbank/best_model/',
                 use_gpu=True, state_dict=None):
        """reference: predict.py:25-67.  ``state_dict`` (extension): weights given in memory instead of a file."""
        if not use_gpu:
            raise _lib.PPVError('use_gpu=False: the B200 build of ppvector has no CPU path')
        assert torch.cuda.is_available(), 'GPU不可用'
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.threshold = threshold
        if isinstance(configs, str):
            with open(configs, 'r', encoding='utf-8') as f:
                configs = yaml.load(f.read(), Loader=yaml.FullLoader)
            print_arguments(configs=configs)
        self.configs = dict_to_object(configs)
        self._audio_featurizer = AudioFeaturizer(feature_method=self.configs.preprocess_conf.feature_method,
                                                 method_args=self.configs.preprocess_conf.get('method_args', {}))
        backbone = build_model(input_size=self._audio_featurizer.feature_dim, configs=self.configs)
        if state_dict is None:
            if not os.path.exists(model_path):
                raise Exception("模型文件不存在，请检查{}是否存在！".format(model_path))
            state_dict = load_state_dict_file(model_path)
        # the reference wraps the backbone in nn.Sequential, hence the "0." key prefix (predict.py:59, checkpoint.py)
        state_dict = {(k[2:] if k.startswith('0.') else k): v for k, v in state_dict.items() if not k.startswith('1.')}
        backbone.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()})
        self.predictor = backbone.eval().to(self.device)
        self._pinned = None
        self._pinned_out = None
        if audio_db_path is not None:
            logger.warning('audio_db_path: the enrolment database is out of scope of the B200 hot path (ignored)')

    # ---- audio loading: predict.py:189-216 ----------------------------------------------------------------
    def _load_audio(self, audio_data, sample_rate=16000):
        if isinstance(audio_data, str):
            audio_segment = AudioSegment.from_file(audio_data)
        elif isinstance(audio_data, BufferedReader):
            audio_segment = AudioSegment.from_file(audio_data)
        elif isinstance(audio_data, np.ndarray):
            audio_segment = AudioSegment.from_ndarray(audio_data, sample_rate)
        elif isinstance(audio_data, bytes):
            audio_segment = AudioSegment.from_bytes(audio_data)
        elif isinstance(audio_data, AudioSegment):
            audio_segment = audio_data
        else:
            raise Exception(f'不支持该数据类型，当前数据类型为：{type(audio_data)}')
        ds = self.configs.dataset_conf.dataset
        assert audio_segment.duration >= ds.min_duration, \
            f'音频太短，最小应该为{ds.min_duration}s，当前音频为{audio_segment.duration}s'
        if audio_segment.sample_rate != ds.sample_rate:
            audio_segment.resample(ds.sample_rate)
        if ds.use_dB_normalization:
            audio_segment.normalize(target_db=ds.target_dB)
        return audio_segment

    # ---- host buffers -> embeddings -----------------------------------------------------------------------
    def extract_embeddings(self, waveforms: np.ndarray, input_lens_ratio=None) -> np.ndarray:
        """[B,L] float32 host waveforms (already normalised / padded) -> [B,embd] float32 host embeddings.
        One pinned H2D copy, one fused library call, one D2H copy."""
        waveforms = np.ascontiguousarray(waveforms, dtype=np.float32)
        B, L = waveforms.shape
        if self._pinned is None or self._pinned.shape[0] < B or self._pinned.shape[1] != L:
            self._pinned = torch.empty((B, L), dtype=torch.float32).pin_memory()
        stage = self._pinned[:B]
        stage.copy_(torch.from_numpy(waveforms))
        return self.extract_embeddings_pinned(stage, input_lens_ratio).numpy().copy()

    def extract_embeddings_pinned(self, stage: torch.Tensor, input_lens_ratio=None) -> torch.Tensor:
        """As extract_embeddings, for a caller that already owns a pinned [B,L] float32 host tensor; returns a
        pinned host tensor (valid until the next call)."""
        B = stage.shape[0]
        wav = stage.to(self.device, non_blocking=True)
        emb = self.predictor.forward_wav(self._audio_featurizer, wav, input_lens_ratio)
        if self._pinned_out is None or self._pinned_out.shape[0] < B:
            self._pinned_out = torch.empty((B, emb.shape[1]), dtype=torch.float32).pin_memory()
        out = self._pinned_out[:B]
        out.copy_(emb, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out

    def extract_embeddings_stream(self, pinned_batches, input_lens_ratio=None):
        """Pipelined form of extract_embeddings_pinned for a sequence of pinned [B,L] float32 host batches: the H2D
        copy of batch i+1 runs on a copy stream while batch i is in the kernels (two device staging buffers, events
        in both directions).  Yields one pinned host [B,embd] tensor per batch, in order; every batch still pays its
        own H2D and D2H -- they are overlapped, not skipped."""
        comp = torch.cuda.current_stream(self.device)
        copy = getattr(self, '_copy_stream', None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(self.device)
        dev = [None, None]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        outs = [None, None]
        it = iter(pinned_batches)

        def stage(i, host):
            k = i & 1
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(free[k])  # the kernels of batch i-2 have consumed this buffer
                if dev[k] is None or dev[k].shape != host.shape:
                    dev[k] = torch.empty(host.shape, dtype=torch.float32, device=self.device)
                dev[k].copy_(host, non_blocking=True)
                ready[k].record(copy)

        nxt = next(it, None)
        if nxt is None:
            return
        stage(0, nxt)
        i = 0
        while nxt is not None:
            k = i & 1
            comp.wait_event(ready[k])
            emb = self.predictor.forward_wav(self._audio_featurizer, dev[k], input_lens_ratio)
            free[k].record(comp)
            if outs[k] is None or outs[k].shape != emb.shape:
                outs[k] = torch.empty(emb.shape, dtype=torch.float32).pin_memory()
            outs[k].copy_(emb, non_blocking=True)
            done[k].record(comp)
            nxt = next(it, None)
            if nxt is not None:
                stage(i + 1, nxt)  # overlaps with the kernels just enqueued
            done[k].synchronize()
            yield outs[k]
            i += 1

    def predict(self, audio_data, sample_rate=16000):
        """reference: predict.py:218-233 -> [embd] numpy"""
        seg = self._load_audio(audio_data=audio_data, sample_rate=sample_rate)
        return self.extract_embeddings(seg.samples[None, :])[0]

    def predict_batch(self, audios_data, sample_rate=16000, batch_size=32):
        """reference: predict.py:235-269.  Zero-pad to the longest of the WHOLE list, lens-ratio mask after CMN (featurizer.py:48-59),
        then the model in chunks of ``batch_size`` (predict.py:264-267).  Featurisation is per utterance, so chunking the padded
        waveforms gives the same embeddings as featurising the whole list at once, and bounds the workspace."""
        segs = [self._load_audio(audio_data=a, sample_rate=sample_rate).samples for a in audios_data]
        max_len = max(s.shape[0] for s in segs)
        inputs = np.zeros((len(segs), max_len), dtype=np.float32)
        ratio = np.zeros((len(segs),), dtype=np.float32)
        for i, s in enumerate(segs):
            inputs[i, :s.shape[0]] = s
            ratio[i] = s.shape[0] / max_len
        batch_size = max(1, int(batch_size))
        return np.concatenate([self.extract_embeddings(inputs[i:i + batch_size], ratio[i:i + batch_size])
                               for i in range(0, len(segs), batch_size)], axis=0)

    def contrast(self, audio_data1, audio_data2):
        """reference: predict.py:271-283 -> cosine similarity of the two embeddings"""
        feature1 = self.predict(audio_data1)
        feature2 = self.predict(audio_data2)
        e = torch.from_numpy(np.stack([feature1, feature2])).to(self.device)
        return float(cosine_matrix(e[:1], e[1:])[0, 0].item())
