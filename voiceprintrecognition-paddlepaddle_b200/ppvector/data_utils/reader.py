"""PPVectorDataset -- the eval / extract_feature / plain-train subset of ppvector/data_utils/reader.py:16-163.

List file lines are ``path\\tlabel``; items are ``(feature [T,F] float32 CUDA tensor, speaker id)``.  Audio goes
wav -> float32 -> (resample) -> dB normalise -> crop (eval: from 0; train: random start; extract_feature: no crop) ->
``AudioFeaturizer`` on the GPU -> SpecAugment (train mode, reader.py:105-107, ``ppv_spec_augment``).  ``.npy`` entries are
pre-extracted features (reader.py:78-83).  Waveform augmentation (speed / volume / noise / reverb, yeaudio) is not implemented
on the B200 path: an ``aug_conf`` that enables any of them raises instead of silently skipping it."""
import random

import numpy as np
import torch
from tqdm import tqdm

from ppvector.data_utils.audio import AudioSegment
from ppvector.data_utils.featurizer import AudioFeaturizer
from ppvector.data_utils.spec_aug import SpecAugmentor


class PPVectorDataset(torch.utils.data.Dataset):
    def __init__(self, data_list_path, audio_featurizer: AudioFeaturizer, max_duration=3, min_duration=0.5, mode='train',
                 sample_rate=16000, aug_conf=None, num_speakers=None, use_dB_normalization=True, target_dB=-20,
                 device='cuda'):
        super().__init__()
        assert mode in ['train', 'eval', 'extract_feature']
        self.spec_augment = None
        if mode == 'train' and aug_conf is not None:
            conf = dict(aug_conf) if isinstance(aug_conf, dict) else dict(vars(aug_conf))
            for name in ('speed', 'volume', 'noise', 'reverb'):
                sub = conf.get(name)
                prob = (sub.get('prob', 0) if isinstance(sub, dict) else getattr(sub, 'prob', 0)) if sub is not None else 0
                if prob and prob > 0:
                    raise NotImplementedError(f'{name} augmentation is not implemented on the B200 path (SURVEY.md §2 row 13); set its prob to 0')
            sa = conf.get('spec_aug')
            if sa is not None:  # reader.py:150-151
                self.spec_augment = SpecAugmentor(**(dict(sa) if isinstance(sa, dict) else dict(vars(sa))))
        self.data_list_path = data_list_path
        self.max_duration, self.min_duration, self.mode = max_duration, min_duration, mode
        self._target_sample_rate = sample_rate
        self._use_dB_normalization, self._target_dB = use_dB_normalization, target_dB
        self.num_speakers = num_speakers
        self.audio_featurizer = audio_featurizer
        self.device = torch.device(device)
        # frame count of a max_duration crop (reader.py:115-119 featurises random noise just to learn T)
        self.max_feature_len = audio_featurizer.num_frames(int(self.max_duration * self._target_sample_rate))
        with open(self.data_list_path, 'r', encoding='utf-8') as f:
            self.lines = [ln for ln in f.readlines() if ln.strip()]
        self.labels = [np.int64(line.strip().split('\t')[1]) for line in self.lines]
        if self.mode == 'eval':
            self.sort_list()

    def load_samples(self, idx):
        """Returns (float32 samples, label) after resample / normalise / crop, or None for a .npy feature entry."""
        data_path, spk_id = self.lines[idx].strip().split('\t')
        if data_path.endswith('.npy'):
            return None, int(spk_id)
        seg = AudioSegment.from_file(data_path)
        if self.mode in ('train', 'extract_feature') and seg.duration < self.min_duration:
            return self.load_samples(idx + 1 if idx < len(self.lines) - 1 else 0)  # reader.py:88-89
        if seg.sample_rate != self._target_sample_rate:
            seg.resample(self._target_sample_rate)
        if self._use_dB_normalization:
            seg.normalize(target_db=self._target_dB)
        x = seg.samples
        if self.mode != 'extract_feature' and seg.duration > self.max_duration:
            n = int(self.max_duration * self._target_sample_rate)
            start = random.randint(0, x.shape[0] - n) if self.mode == 'train' else 0
            x = x[start:start + n]
        return x, int(spk_id)

    def __getitem__(self, idx):
        x, spk_id = self.load_samples(idx)
        if x is None:
            data_path = self.lines[idx].strip().split('\t')[0]
            feature = np.load(data_path)
            if feature.shape[0] > self.max_feature_len:
                s = random.randint(0, feature.shape[0] - self.max_feature_len) if self.mode == 'train' else 0
                feature = feature[s:s + self.max_feature_len, :]
            feature = torch.from_numpy(feature.astype(np.float32)).to(self.device)
        else:
            feature = self.audio_featurizer(torch.from_numpy(x).to(self.device)).squeeze(0)
        if self.mode == 'train' and self.spec_augment is not None:  # reader.py:105-107
            feature = self.spec_augment(feature)
        return feature, spk_id

    def __len__(self):
        return len(self.lines)

    def sort_list(self):
        """reader.py:122-138: eval lists are processed in order of increasing duration"""
        lengths = []
        for line in tqdm(self.lines, desc=f"对列表[{self.data_list_path}]进行长度排序"):
            data_path = line.split('\t')[0]
            if data_path.endswith('.npy'):
                lengths.append(np.load(data_path, mmap_mode='r').shape[0])
            else:
                lengths.append(AudioSegment.from_file(data_path).duration)
        order = np.argsort(lengths, kind='stable')
        self.lines = [self.lines[i] for i in order]
        self.labels = [self.labels[i] for i in order]
