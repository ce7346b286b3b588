"""build_model -- reference: ppvector/models/__init__.py:15-21 (string-named factory)."""
import importlib

from loguru import logger

from .campplus import CAMPPlus
from .ecapa_tdnn import EcapaTdnn
from .eres2net import ERes2Net
from .resnet_se import ResNetSE

__all__ = ['build_model']

# Models of the reference that the B200 path does not implement yet (ERes2NetV2 and the plain Res2Net / TDNN variants, "next").
_NOT_YET = ('ERes2NetV2', 'Res2Net', 'TDNN')


def build_model(input_size, configs):
    use_model = configs.model_conf.get('model', 'CAMPPlus')
    model_args = configs.model_conf.get('model_args', {})
    if use_model in _NOT_YET:
        raise NotImplementedError(f'{use_model} 尚未在 B200 路径实现 (EcapaTdnn, ResNetSE, ERes2Net and CAMPPlus are implemented; no fallback)')
    mod = importlib.import_module(__name__)
    model = getattr(mod, use_model)(input_size=input_size, **model_args)
    logger.info(f'成功创建模型：{use_model}，参数为：{model_args}')
    return model
