"""GPU: the reference-facing Python surface -- PPVectorPredictor.predict / predict_batch / contrast (predict.py:218-283),
PPVectorTrainer.extract_features / evaluate (trainer.py:134-157, 367-447) -- on wav files written from the golden PCM of
the reference's bundled audio, against the oracle."""
import os
import wave

import numpy as np
import pytest
import torch
import yaml

from oracle import ecapa as oe
from oracle import fbank as ofb
from oracle import head as oh
from ppvector.metric.metrics import compute_dcf, compute_eer, compute_fnr_fpr
from ppvector.predict import PPVectorPredictor
from ppvector.trainer import PPVectorTrainer

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["a_1", "a_2", "b_1", "b_2", "long3s"]


@pytest.fixture(scope="module")
def wavs(tmp_path_factory, golden_dir):
    d = tmp_path_factory.mktemp("wavs")
    g = np.load(f"{golden_dir}/fbank_wavs.npz")
    paths = {}
    for n in NAMES:
        p = str(d / f"{n}.wav")
        with wave.open(p, "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(16000)
            w.writeframes(g[n + "_pcm"].astype("<i2").tobytes())
        paths[n] = p
    return paths, g


@pytest.fixture(scope="module")
def cfg():
    return yaml.load(open(os.path.join(ROOT, "configs", "ecapa_tdnn.yml")), Loader=yaml.FullLoader)


@pytest.fixture(scope="module")
def W64():
    return oe.make_ecapa_weights(seed=1000, dtype=torch.float64)


def oracle_embed(pcm, W64, max_samples=None):
    x = ofb.db_normalize(pcm.astype(np.float32) / 32768.0, -20.0)
    if max_samples is not None:
        x = x[:max_samples]
    feat = torch.from_numpy(ofb.audio_featurizer_fbank(x, None, dtype=np.float64, n_mels=80))
    return oe.ecapa_forward(feat, W64)[0].numpy()


def test_predictor_predict_batch_contrast(cuda, wavs, cfg, W64):
    paths, g = wavs
    sd = {k: v.float().numpy() for k, v in W64.items()}
    pred = PPVectorPredictor(cfg, model_path=None, use_gpu=True, state_dict=sd)
    ref = {n: oracle_embed(g[n + "_pcm"], W64) for n in NAMES}
    e1 = pred.predict(paths["a_1"])
    assert e1.shape == (192,) and e1.dtype == np.float32
    assert 1 - oh.cosine_pair(e1, ref["a_1"]) < 1e-8
    # numpy input and bytes input go through the same path (predict.py:196-205)
    e1b = pred.predict(g["a_1_pcm"], sample_rate=16000)
    assert np.abs(e1 - e1b).max() < 1e-6
    e1c = pred.predict(open(paths["a_1"], "rb").read())
    assert np.abs(e1 - e1c).max() < 1e-6
    # contrast == cosine of the two oracle embeddings
    c = pred.contrast(paths["a_1"], paths["b_2"])
    assert abs(c - oh.cosine_pair(ref["a_1"], ref["b_2"])) < 1e-4
    # predict_batch: zero-padded waveforms + lens ratio (CMN over padded frames, mask after): compare with the oracle
    # run the same way
    batch = pred.predict_batch([paths[n] for n in NAMES])
    assert batch.shape == (5, 192)
    segs = [ofb.db_normalize(g[n + "_pcm"].astype(np.float32) / 32768.0, -20.0) for n in NAMES]
    L = max(len(s) for s in segs)
    x = np.zeros((5, L), np.float32)
    for i, s in enumerate(segs):
        x[i, :len(s)] = s
    ratio = np.array([len(s) / L for s in segs], np.float32)
    feat = torch.from_numpy(ofb.audio_featurizer_fbank(x, ratio, dtype=np.float64, n_mels=80))
    refb = oe.ecapa_forward(feat, W64).numpy()
    assert np.abs(oh.cosine_matrix(batch, batch) - oh.cosine_matrix(refb, refb)).max() < 1e-4
    with pytest.raises(AssertionError):
        pred.predict(np.zeros(1000, np.float32))  # shorter than min_duration (predict.py:207-209)
    from ppvector._lib import PPVError
    with pytest.raises(PPVError):
        PPVectorPredictor(cfg, model_path=None, use_gpu=False, state_dict=sd)


def test_trainer_extract_features_and_evaluate(cuda, wavs, cfg, W64, tmp_path):
    paths, g = wavs
    import copy
    cfg = copy.deepcopy(cfg)
    lists = {}
    spk = {"a_1": 0, "a_2": 0, "b_1": 1, "b_2": 1, "long3s": 2}
    for name, members in {"train": NAMES, "enroll": ["a_1", "b_1", "long3s"], "trials": ["a_2", "b_2"]}.items():
        p = str(tmp_path / f"{name}_list.txt")
        with open(p, "w") as f:
            for n in members:
                f.write(f"{paths[n]}\t{spk[n]}\n")
        lists[name] = p
    cfg["dataset_conf"]["train_list"], cfg["dataset_conf"]["enroll_list"], cfg["dataset_conf"]["trials_list"] = \
        lists["train"], lists["enroll"], lists["trials"]
    sd = {k: v.float().numpy() for k, v in W64.items()}
    tr = PPVectorTrainer(cfg, use_gpu=True, state_dict=sd)
    # config 1 of BASELINE.json: Fbank-80 extraction of the bundled wavs through the data_utils surface
    tr.extract_features(save_dir=str(tmp_path / "features"), max_duration=100)
    out_list = lists["train"].replace(".txt", "_features.txt")
    lines = open(out_list).read().strip().split("\n")
    assert len(lines) == 5
    frames = {"a_1": 365, "a_2": 218, "b_1": 502, "b_2": 516, "long3s": 298}
    for line, n in zip(lines, NAMES):
        path, label = line.split("\t")
        feat = np.load(path)
        assert feat.shape == (frames[n], 80) and int(label) == spk[n]
        x = ofb.db_normalize(g[n + "_pcm"].astype(np.float32) / 32768.0, -20.0)
        ref = ofb.audio_featurizer_fbank(x, None, dtype=np.float64, n_mels=80)[0]
        assert np.abs(feat - ref).max() < 2e-3
    # evaluate: trial x enrol cosine matrix -> EER / minDCF, against the oracle embeddings scored the same way
    eer, min_dcf, threshold = tr.evaluate()
    # oracle run the way the reference evaluates (trainer.py:391-410 + collate_fn.py): lists sorted by duration, features
    # zero-padded to the longest of the batch, NO length mask reaching the model (quirk kept)
    def oracle_list(members):
        members = sorted(members, key=lambda n: len(g[n + "_pcm"]))
        feats = [ofb.audio_featurizer_fbank(ofb.db_normalize(g[n + "_pcm"].astype(np.float32) / 32768.0, -20.0), None,
                                            dtype=np.float64, n_mels=80)[0] for n in members]
        Tm_ = max(f.shape[0] for f in feats)
        pad = np.zeros((len(feats), Tm_, 80))
        for i, f in enumerate(feats):
            pad[i, :f.shape[0]] = f
        return oe.ecapa_forward(torch.from_numpy(pad), W64).numpy(), np.array([spk[n] for n in members])
    E, e_lab = oracle_list(["a_1", "b_1", "long3s"])
    Tm, t_lab = oracle_list(["a_2", "b_2"])
    scores = oh.cosine_matrix(Tm, E).astype(np.float32).reshape(-1)
    labels = (t_lab[:, None] == e_lab[None, :]).astype(np.int32).reshape(-1)
    fnr, fpr, _ = compute_fnr_fpr(scores, labels)
    eer_ref, thr_ref = compute_eer(fnr, fpr, scores)
    assert abs(eer - float(eer_ref)) < 1e-6 and abs(threshold - float(thr_ref)) < 1e-4
    assert abs(min_dcf - float(compute_dcf(fnr, fpr))) < 1e-6


@pytest.mark.parametrize("model_name", ["CAMPPlus", "ERes2Net", "ERes2NetV2", "ResNetSE"])
def test_predictor_with_the_other_backbones(cuda, wavs, cfg, model_name):
    """PPVectorPredictor (predict.py:218-283) is model-agnostic in the reference: the same surface must drive every backbone."""
    import copy
    import importlib
    paths, g = wavs
    cfg = copy.deepcopy(cfg)
    cfg["model_conf"]["model"] = model_name
    cfg["model_conf"]["model_args"] = {"embd_dim": 192}
    import functools
    om = importlib.import_module({"CAMPPlus": "oracle.campplus", "ERes2Net": "oracle.eres2net", "ERes2NetV2": "oracle.eres2net",
                                  "ResNetSE": "oracle.resnet_se"}[model_name])
    make = {"CAMPPlus": "make_campplus_weights", "ERes2Net": "make_eres2net_weights", "ERes2NetV2": "make_eres2net_weights",
            "ResNetSE": "make_resnet_se_weights"}[model_name]
    fwd = {"CAMPPlus": "campplus_forward", "ERes2Net": "eres2net_forward", "ERes2NetV2": "eres2net_forward", "ResNetSE": "resnet_se_forward"}[model_name]
    v2 = {"base_width": 26, "version": 2} if model_name == "ERes2NetV2" else {}
    Wm = getattr(om, make)(seed=1000, dtype=torch.float64, **v2)
    om = type("O", (), {fwd: staticmethod(functools.partial(getattr(om, fwd), **v2))})
    p = PPVectorPredictor(configs=cfg, state_dict={k: v.float().numpy() for k, v in Wm.items()})
    emb = p.predict(paths["a_2"])
    x = ofb.db_normalize(g["a_2_pcm"].astype(np.float32) / 32768.0, -20.0)
    feat = torch.from_numpy(ofb.audio_featurizer_fbank(x, None, dtype=np.float64, n_mels=80))
    ref = getattr(om, fwd)(feat, Wm)[0].numpy()
    cos = float((emb * ref).sum() / np.linalg.norm(emb) / np.linalg.norm(ref))
    assert emb.shape == (192,) and 1 - cos < 1e-6, cos
    s = p.contrast(paths["a_2"], paths["b_2"])
    e2 = getattr(om, fwd)(torch.from_numpy(ofb.audio_featurizer_fbank(ofb.db_normalize(g["b_2_pcm"].astype(np.float32) / 32768.0, -20.0), None,
                                                                      dtype=np.float64, n_mels=80)), Wm)[0].numpy()
    assert abs(s - float((ref * e2).sum() / np.linalg.norm(ref) / np.linalg.norm(e2))) < 1e-4


def test_trainer_train_runs_the_cuda_step(cuda, wavs, cfg, tmp_path):
    """PPVectorTrainer.train (trainer.py:281-365): list file -> features -> SpecAugment -> CUDA training step -> Adam with the
    reference's LR / margin schedules -> checkpoint with the reference's key names -> evaluate on it."""
    import copy
    paths, g = wavs
    cfg = copy.deepcopy(cfg)
    spk = {"a_1": 0, "a_2": 0, "b_1": 1, "b_2": 1, "long3s": 2}
    lists = {}
    for name, members in {"train": NAMES + NAMES, "enroll": ["a_1", "b_1", "long3s"], "trials": ["a_2", "b_2"]}.items():
        p = str(tmp_path / f"{name}_list.txt")
        with open(p, "w") as f:
            for n in members:
                f.write(f"{paths[n]}\t{spk[n]}\n")
        lists[name] = p
    cfg["dataset_conf"]["train_list"], cfg["dataset_conf"]["enroll_list"], cfg["dataset_conf"]["trials_list"] = \
        lists["train"], lists["enroll"], lists["trials"]
    cfg["dataset_conf"]["sampler"]["batch_size"] = 4
    cfg["model_conf"]["classifier"]["num_speakers"] = 3
    cfg["train_conf"]["max_epoch"] = 2
    cfg["train_conf"]["log_interval"] = 1
    aug = {"spec_aug": {"prob": 0.5, "freq_mask_ratio": 0.1, "n_freq_masks": 1, "time_mask_ratio": 0.05, "n_time_masks": 1, "max_time_warp": 0},
           "noise": {"prob": 0.0}}
    tr = PPVectorTrainer(cfg, use_gpu=True, data_augment_configs=aug)
    save = str(tmp_path / "models")
    history = tr.train(save_model_path=save, do_eval=True)
    assert len(history) == 4 and all(np.isfinite(history)) and history[0] > 0.5  # 10 utterances / batch 4, drop_last: 2 steps x 2 epochs
    assert tr.engine.step_count == 4 and float(tr.engine.exp_avg_sq.abs().sum()) > 0
    # the reference's checkpoint layout (utils/checkpoint.py:104-159): <model>_<feature>/{epoch_N, last_model, best_model}
    import json
    root = os.path.join(save, "EcapaTdnn_Fbank")
    assert sorted(os.listdir(root)) == ["best_model", "epoch_1", "epoch_2", "last_model"]
    ck = torch.load(os.path.join(root, "last_model", "model.pt"))
    assert "0.blocks.0.conv.conv.weight" in ck and ck["1.weight"].shape == (192, 3)
    state = json.load(open(os.path.join(root, "last_model", "model.state")))
    assert state["last_epoch"] == 2 and state["model_conf.model"] == "EcapaTdnn" and "eer" in state and "margin" in state
    opt = torch.load(os.path.join(root, "last_model", "optimizer.pt"))
    assert opt["step_count"] == 4 and torch.equal(opt["exp_avg_sq"], tr.engine.exp_avg_sq.cpu())
    eer, min_dcf, thr = PPVectorTrainer(cfg, use_gpu=True).evaluate(resume_model=os.path.join(root, "best_model"))
    assert 0.0 <= eer <= 1.0 and np.isfinite(thr)
    # resume: a third epoch continues from last_model -- Adam moments, step count, LR / margin schedule positions and the epoch
    cfg3 = copy.deepcopy(cfg)
    cfg3["train_conf"]["max_epoch"] = 3
    tr3 = PPVectorTrainer(cfg3, use_gpu=True, data_augment_configs=aug)
    h3 = tr3.train(save_model_path=save, do_eval=False)
    assert len(h3) == 2 and tr3.engine.step_count == 6 and tr3.train_step == 6
    assert json.load(open(os.path.join(root, "last_model", "model.state")))["last_epoch"] == 3
    assert not os.path.exists(os.path.join(root, "epoch_0")) and os.path.exists(os.path.join(root, "epoch_3"))
    cfg2 = copy.deepcopy(cfg)
    cfg2["model_conf"]["model"] = "ResNetSE"
    with pytest.raises(NotImplementedError):
        PPVectorTrainer(cfg2, use_gpu=True).train()


@pytest.mark.parametrize("lanes", [1, 2, 3])
def test_streaming_lanes_equal_the_single_call(cuda, cfg, W64, lanes):
    """extract_embeddings_stream / embed_resident_stream keep `lanes` batches in the kernels at once (replica models on their own streams):
    every batch must come back in order and bitwise equal to the one-batch-at-a-time call, for more batches than lanes and staging buffers."""
    pred = PPVectorPredictor(cfg, model_path=None, use_gpu=True, state_dict={k: v.float().numpy() for k, v in W64.items()})
    g = torch.Generator().manual_seed(31)
    host = [(0.1 * torch.randn(5, 16000, generator=g)).pin_memory() for i in range(7)]
    want = [pred.extract_embeddings(h.numpy()) for h in host]
    got = [out.numpy().copy() for out in pred.extract_embeddings_stream(iter(host), lanes=lanes)]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    dev = [h.to(cuda) for h in host]
    n = 0
    for a, b in zip(pred.embed_resident_stream(dev, lanes=lanes), want):  # a generator: each embedding is complete when it is yielded
        assert np.array_equal(a.cpu().numpy(), b)
        n += 1
    assert n == len(want)
