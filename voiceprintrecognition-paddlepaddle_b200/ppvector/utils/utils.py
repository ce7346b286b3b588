"""CLI / config helpers -- same semantics as ppvector/utils/utils.py:8-52 of the reference
(``add_arguments`` with strtobool for bools, ``dict_to_object`` attr-dict, ``print_arguments``)."""
from loguru import logger


def _strtobool(v):
    v = str(v).lower()
    if v in ('y', 'yes', 't', 'true', 'on', '1'):
        return True
    if v in ('n', 'no', 'f', 'false', 'off', '0'):
        return False
    raise ValueError(f'invalid truth value {v!r}')


def add_arguments(argname, type, default, help, argparser, **kwargs):
    """reference: utils.py:32-38 (distutils.util.strtobool for bool arguments)"""
    type = _strtobool if type == bool else type
    argparser.add_argument("--" + argname, default=default, type=type, help=help + ' 默认: %(default)s.', **kwargs)


class Dict(dict):
    """reference: utils.py:41-43"""
    __setattr__ = dict.__setitem__
    __getattr__ = dict.__getitem__


def dict_to_object(dict_obj):
    """reference: utils.py:46-52"""
    if not isinstance(dict_obj, dict):
        return dict_obj
    inst = Dict()
    for k, v in dict_obj.items():
        inst[k] = dict_to_object(v)
    return inst


def print_arguments(args=None, configs=None, title=None):
    """reference: utils.py:8-29"""
    if args:
        logger.info("----------- 额外配置参数 -----------")
        for arg, value in sorted(vars(args).items()):
            logger.info(f"{arg}: {value}")
    if configs:
        logger.info(f"----------- {title or '配置文件参数'} -----------")

        def walk(d, depth):
            for k, v in sorted(d.items()):
                if isinstance(v, dict):
                    logger.info("\t" * depth + f"{k}:")
                    walk(v, depth + 1)
                else:
                    logger.info("\t" * depth + f"{k}: {v}")
        walk(configs, 0)


def cosin_metric(x1, x2):
    """reference: utils.py:82-83 -- a single pair is host scalar math; batches go through ppvector.metric.cosine"""
    import numpy as np
    return float(np.dot(x1, x2) / (np.linalg.norm(x1) * np.linalg.norm(x2)))


def cal_accuracy(y_score, y_true, threshold=0.5):
    """Fraction of trials decided correctly when scores >= threshold count as 'same speaker' (reference: utils.py:73-79)."""
    import numpy as np
    decided = np.asarray(y_score) >= threshold
    return float(np.mean(decided == np.asarray(y_true).astype(bool)))


def cal_accuracy_threshold(y_score, y_true):
    """Best accuracy over the thresholds 0.00, 0.01, ..., 0.99 and the first threshold reaching it (reference: utils.py:56-69);
    one vectorised comparison [100, n] instead of a Python loop."""
    import numpy as np
    scores, truth = np.asarray(y_score), np.asarray(y_true).astype(bool)
    grid = np.arange(100) * 0.01
    acc = ((scores[None, :] >= grid[:, None]) == truth[None, :]).mean(axis=1)
    best = int(np.argmax(acc))  # argmax returns the first maximum, like the reference's strict '>' update
    if acc[best] <= 0:
        return 0, 0
    return float(acc[best]), float(grid[best])
