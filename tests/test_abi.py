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
