"""build_loss -- reference: ppvector/loss/__init__.py:16-22."""
import importlib

from loguru import logger

from .aamloss import AAMLoss

__all__ = ['build_loss']


def build_loss(configs):
    use_loss = configs.loss_conf.get('loss', 'AAMLoss')
    loss_args = configs.loss_conf.get('loss_args', {})
    if use_loss != 'AAMLoss':
        raise NotImplementedError(f'{use_loss} 尚未在 B200 路径实现 (only AAMLoss is implemented)')
    mod = importlib.import_module(__name__)
    loss = getattr(mod, use_loss)(**loss_args)
    logger.info(f'成功创建损失函数：{use_loss}，参数为：{loss_args}')
    return loss
