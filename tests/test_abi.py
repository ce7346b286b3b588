"""CPU: the C-ABI library loads and exports every symbol include/ppv_b200.h declares; the ctypes table in
ppvector/_lib.py covers the same set; the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from ppvector import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ppv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = header_functions()
    assert len(names) >= 25
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ppv_b200.h but not exported by libppv_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in ppvector/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_buffer():
    lib = _lib.load()
    assert lib.ppv_version() == 100
    buf = C.create_string_buffer(8)
    assert lib.ppv_last_error(buf, 8) >= 0


def test_default_configs_match_reference_yaml():
    lib = _lib.load()
    f = _lib.FbankCfg()
    lib.ppv_fbank_default_cfg(C.byref(f))
    assert (f.sample_rate, f.n_mels, f.frame_length_ms, f.frame_shift_ms) == (16000, 80, 25.0, 10.0)
    assert abs(f.preemph - 0.97) < 1e-7 and f.low_freq == 20.0 and f.high_freq == 0.0
    e = _lib.EcapaCfg()
    lib.ppv_ecapa_default_cfg(C.byref(e))
    # configs/ecapa_tdnn.yml:50-59 + ecapa_tdnn.py:151-158 defaults
    assert list(e.channels) == [512, 512, 512, 512, 1536] and list(e.kernel_sizes) == [5, 3, 3, 3, 1]
    assert list(e.dilations) == [1, 2, 3, 4, 1] and e.embd_dim == 192 and e.input_size == 80


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    f = _lib.FbankCfg()
    lib.ppv_fbank_default_cfg(C.byref(f))
    h = C.c_void_p()
    assert lib.ppv_fbank_create(C.byref(f), C.byref(h)) != 0
    assert "no CPU fallback" in _lib.last_error()
    from ppvector.data_utils.featurizer import AudioFeaturizer
    fz = AudioFeaturizer("Fbank", {"sr": 16000, "n_mels": 80})
    assert fz.feature_dim == 80
    with pytest.raises(_lib.PPVError):
        fz(torch.zeros(1, 16000))  # CPU tensor -> loud failure


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "voiceprintrecognition-paddlepaddle_b200")
    for d, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, fn), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(d, fn)


def test_every_shipped_config_builds_its_backbone():
    """configs/*.yml keep the reference's schema; build_model (models/__init__.py:15-21) must construct each backbone from it
    (parameter holders only: no GPU, no library call)."""
    import glob
    import os

    import yaml

    from ppvector.models import build_model
    from ppvector.utils.utils import dict_to_object
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seen = {}
    for path in sorted(glob.glob(os.path.join(root, "configs", "*.yml"))):
        if os.path.basename(path) == "augmentation.yml":  # the augmentation schema, not a model config
            aug = yaml.load(open(path), Loader=yaml.FullLoader)
            assert set(aug) == {"speed", "volume", "noise", "reverb", "spec_aug"}
            continue
        cfg = dict_to_object(yaml.load(open(path), Loader=yaml.FullLoader))
        for key in ("dataset_conf", "preprocess_conf", "model_conf", "loss_conf", "optimizer_conf", "train_conf"):
            assert key in cfg, (path, key)
        m = build_model(input_size=80, configs=cfg)
        seen[cfg.model_conf.model] = sum(p.numel() for p in m.parameters())
    assert set(seen) == {"EcapaTdnn", "ResNetSE", "ERes2Net", "CAMPPlus"}
    assert seen["EcapaTdnn"] == 6194048 and seen["ERes2Net"] == 6620128 and seen["CAMPPlus"] == 6859232


def test_loss_registry_and_head_selectors():
    """loss/__init__.py:16-22 of the reference resolves seven loss classes by name; six run on the fused CUDA head here.  The head selector the
    training loop forwards to the C ABI (include/ppv_b200.h: PPV_HEAD_*) is checked bit by bit; the seventh raises by name."""
    from ppvector import _lib
    from ppvector.loss import build_loss
    from ppvector.utils.utils import dict_to_object

    def mk(name, **args):
        return build_loss(dict_to_object({"loss_conf": {"loss": name, "loss_args": args}}))

    assert mk("AAMLoss", margin=0.2, scale=32).easy_margin in (0, False)
    assert int(mk("AAMLoss", margin=0.2, scale=32, easy_margin=True).easy_margin) == _lib.PPV_HEAD_AAM_EASY
    assert mk("AMLoss").easy_margin == _lib.PPV_HEAD_AM and mk("ARMLoss").easy_margin == _lib.PPV_HEAD_ARM
    ce = mk("CELoss", label_smoothing=0.1)
    assert ce.easy_margin == _lib.PPV_HEAD_CE and ce.scale == 1.0 and ce.label_smoothing == 0.1
    sub = mk("SubCenterLoss", K=3, margin=0.3)
    assert sub.easy_margin == (_lib.PPV_HEAD_SUBCENTER | (3 << 5)) and sub.margin == 0.3
    assert mk("SubCenterLoss", K=2, easy_margin=True).easy_margin == (_lib.PPV_HEAD_SUBCENTER | (2 << 5) | 1)
    sf = mk("SphereFace2", margin=0.15, lanbuda=0.6, t=3, margin_type="A")
    assert sf.easy_margin == (_lib.PPV_HEAD_SPHEREFACE2 | (3 << 5) | 1) and sf.label_smoothing == 0.6  # lanbuda travels in that slot
    assert mk("SphereFace2").easy_margin == (_lib.PPV_HEAD_SPHEREFACE2 | (3 << 5))
    sub.update(margin=0.25)
    assert sub.margin == 0.25
    with pytest.raises(NotImplementedError, match="TripletAngularMarginLoss"):
        mk("TripletAngularMarginLoss")
    # the selectors do not collide: plain heads 0..4, SphereFace2 has bit 3, SubCenter bit 4, parameters from bit 5 up
    assert _lib.PPV_HEAD_CE < _lib.PPV_HEAD_SPHEREFACE2 < _lib.PPV_HEAD_SUBCENTER < 32


def test_eres2netv2_parameter_names_and_shapes():
    """ERes2NetV2 (reference eres2net.py:379-438): the mirror's state_dict keys and shapes equal the reference class's (the oracle's shape table is
    itself checked against the reference in tests/test_oracle_vs_reference.py by loading its weights into the reference model)."""
    from oracle import eres2net as oe
    from ppvector.models import build_model
    from ppvector.utils.utils import dict_to_object
    m = build_model(input_size=80, configs=dict_to_object({"model_conf": {"model": "ERes2NetV2", "model_args": {"embd_dim": 192}}}))
    S = oe.eres2net_param_shapes(base_width=26, version=2)
    sd = m.state_dict()
    assert sorted(sd) == sorted(S)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(S[k]), k
    assert "layer3_ds.weight" in sd and "fuse34.local_att.0.weight" in sd and "layer1_downsample.weight" not in sd
    assert sd["layer1.0.convs.0.weight"].shape == (13, 13, 3, 3) and sd["layer4.0.convs.1.weight"].shape == (104, 104, 3, 3)
