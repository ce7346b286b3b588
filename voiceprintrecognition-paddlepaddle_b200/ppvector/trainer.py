"""PPVectorTrainer -- the feature-extraction and evaluation halves of ppvector/trainer.py:33-474.

``extract_features`` (trainer.py:134-157) and ``evaluate`` (trainer.py:367-447) keep their signatures, list-file formats and
return values; featurisation, the backbone and the trial x enrol cosine matrix run on the GPU through libppv_b200, EER /
minDCF stay host numpy.  ``train`` needs the backward / optimizer kernels, which are not built
yet: it raises instead of falling back to anything."""
import os

import numpy as np
import torch
import yaml
from loguru import logger
from tqdm import tqdm

from ppvector import _lib
from ppvector.data_utils.collate_fn import collate_fn
from ppvector.data_utils.featurizer import AudioFeaturizer
from ppvector.data_utils.reader import PPVectorDataset
from ppvector.metric.cosine import cosine_matrix
from ppvector.metric.metrics import compute_dcf, compute_eer, compute_fnr_fpr
from ppvector.models import build_model
from ppvector.utils.checkpoint import load_state_dict_file
from ppvector.utils.utils import dict_to_object, print_arguments


class PPVectorTrainer(object):
    def __init__(self, configs, use_gpu=True, data_augment_configs=None, state_dict=None):
        """reference: trainer.py:34-81.  ``state_dict`` (extension): backbone weights given in memory."""
        if not use_gpu:
            raise _lib.PPVError('use_gpu=False: the B200 build of ppvector has no CPU path')
        assert torch.cuda.is_available(), 'GPU不可用'
        self.use_gpu = use_gpu
        self.device = torch.device('cuda', torch.cuda.current_device())
        if isinstance(configs, str):
            with open(configs, 'r', encoding='utf-8') as f:
                configs = yaml.load(f.read(), Loader=yaml.FullLoader)
            print_arguments(configs=configs)
        self.configs = dict_to_object(configs)
        if isinstance(data_augment_configs, str):
            with open(data_augment_configs, 'r', encoding='utf-8') as f:
                data_augment_configs = yaml.load(f.read(), Loader=yaml.FullLoader)
        self.data_augment_configs = dict_to_object(data_augment_configs) if data_augment_configs else None
        self.model = None
        self.audio_featurizer = None
        self.enroll_dataset = self.trials_dataset = None
        self._state_dict = state_dict
        self.stop_train, self.stop_eval = False, False

    def _featurizer(self):
        if self.audio_featurizer is None:
            self.audio_featurizer = AudioFeaturizer(feature_method=self.configs.preprocess_conf.feature_method,
                                                    method_args=self.configs.preprocess_conf.get('method_args', {}))
        return self.audio_featurizer

    # ---- trainer.py:134-157 ------------------------------------------------------------------------------------
    def extract_features(self, save_dir='dataset/features', max_duration=100):
        fz = self._featurizer()
        for list_idx, data_list in enumerate([self.configs.dataset_conf.train_list, self.configs.dataset_conf.enroll_list,
                                              self.configs.dataset_conf.trials_list]):
            if not os.path.exists(data_list):
                logger.warning(f'{data_list} 不存在，跳过')
                continue
            dataset_args = dict(self.configs.dataset_conf.get('dataset', {}))
            dataset_args['max_duration'] = max_duration
            dataset = PPVectorDataset(data_list_path=data_list, audio_featurizer=fz, mode='extract_feature',
                                      device=self.device, **dataset_args)
            save_data_list = data_list.replace('.txt', '_features.txt')
            with open(save_data_list, 'w', encoding='utf-8') as f:
                for i in tqdm(range(len(dataset))):
                    feature, label = dataset[i]
                    # the reference names files by millisecond timestamps (collisions once extraction is fast): use counters
                    save_path = os.path.join(save_dir, str(label), f'{list_idx}_{i:08d}.npy').replace('\\', '/')
                    os.makedirs(os.path.dirname(save_path), exist_ok=True)
                    np.save(save_path, feature.cpu().numpy())
                    f.write(f'{save_path}\t{label}\n')
            logger.info(f'{data_list}列表中的数据已提取特征完成，新列表为：{save_data_list}')

    # ---- model -------------------------------------------------------------------------------------------------
    def _setup_model(self, resume_model=None):
        if self.model is None:
            self.model = build_model(input_size=self._featurizer().feature_dim, configs=self.configs)
        sd = self._state_dict
        if resume_model is not None:
            sd = load_state_dict_file(resume_model)
        if sd is not None:
            # Sequential(backbone, classifier) prefixes of the reference checkpoints: "0." backbone, "1." classifier
            sd = {(k[2:] if k.startswith('0.') else k): v for k, v in sd.items() if not k.startswith('1.')}
            self.model.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()})
        self.model.eval().to(self.device)
        return self.model

    def _embed_list(self, data_list, desc):
        dataset_args = dict(self.configs.dataset_conf.get('dataset', {}))
        dataset_args['max_duration'] = self.configs.dataset_conf.eval_conf.max_duration
        batch_size = self.configs.dataset_conf.eval_conf.batch_size
        dataset = PPVectorDataset(data_list_path=data_list, audio_featurizer=self._featurizer(), mode='eval',
                                  device=self.device, **dataset_args)
        feats, labels = [], []
        for i in tqdm(range(0, len(dataset), batch_size), desc=desc):
            if self.stop_eval:
                break
            batch = [dataset[j] for j in range(i, min(i + batch_size, len(dataset)))]
            features, label, _input_lens = collate_fn(batch)  # input_lens never reaches the model (trainer.py:392-395)
            feats.append(self.model(features))
            labels.append(label)
        return torch.cat(feats, dim=0), torch.cat(labels).numpy().astype(np.int32)

    # ---- trainer.py:367-447 ------------------------------------------------------------------------------------
    def evaluate(self, resume_model=None, save_image_path=None):
        self._setup_model(resume_model)
        with torch.no_grad():
            enroll_features, enroll_labels = self._embed_list(self.configs.dataset_conf.enroll_list, '注册音频声纹特征')
            trials_features, trials_labels = self._embed_list(self.configs.dataset_conf.trials_list, '验证音频声纹特征')
        if self.stop_eval:
            return -1, -1, -1
        # the reference scores one trial against all enrolments per Python iteration (trainer.py:416-423): one GEMM here
        scores = cosine_matrix(trials_features, enroll_features).cpu().numpy().astype(np.float32)
        all_score = scores.reshape(-1)
        all_labels = (trials_labels[:, None] == enroll_labels[None, :]).astype(np.int32).reshape(-1)
        fnr, fpr, thresholds = compute_fnr_fpr(all_score, all_labels)
        eer, threshold = compute_eer(fnr, fpr, all_score)
        min_dcf = compute_dcf(fnr, fpr)
        if save_image_path:
            logger.warning('save_image_path: plotting is out of scope of the B200 hot path (ignored)')
        return float(eer), float(min_dcf), float(threshold)

    def train(self, *args, **kwargs):
        raise NotImplementedError('training (backbone backward, Adam, DDP all-reduce: SURVEY.md §8 row a11) is not built yet; '
                                  'the AAM head forward/backward is available as ppvector.loss.AAMLoss')

    def export(self, *args, **kwargs):
        raise NotImplementedError('export is broken in the reference (trainer.py:467-469) and out of scope')
