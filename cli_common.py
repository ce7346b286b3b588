"""Shared plumbing of the command-line entry points (train.py, eval.py, extract_features.py, infer_contrast.py).

Each script declares its options as a table of (name, type, default, help); the names, types and defaults are the ones the
reference's scripts of the same name accept, so existing command lines keep working (booleans are given as text: ``--use_gpu
False``)."""
import argparse
import os
import sys

PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'voiceprintrecognition-paddlepaddle_b200')
if PKG_DIR not in sys.path:
    sys.path.insert(0, PKG_DIR)

_TRUE, _FALSE = {'1', 'true', 't', 'yes', 'y', 'on'}, {'0', 'false', 'f', 'no', 'n', 'off'}


def text_to_bool(text):
    key = str(text).strip().lower()
    if key in _TRUE:
        return True
    if key in _FALSE:
        return False
    raise argparse.ArgumentTypeError(f'expected a boolean, got {text!r}')


def none_or_str(text):
    return None if text in (None, 'None', 'none', '') else str(text)


def parse_options(description, table, argv=None):
    parser = argparse.ArgumentParser(description=description, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for name, kind, default, help_text in table:
        convert = text_to_bool if kind is bool else none_or_str if kind is str else kind
        parser.add_argument(f'--{name}', type=convert, default=default, help=help_text)
    options = parser.parse_args(argv)
    width = max(len(row[0]) for row in table)
    print('options:')
    for name, *_ in table:
        print(f'  {name:<{width}} = {getattr(options, name)}')
    return options


def init_distributed_if_launched():
    """Under torchrun (WORLD_SIZE > 1): pick this rank's GPU and join the NCCL group."""
    if int(os.environ.get('WORLD_SIZE', '1')) <= 1:
        return False
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group('nccl')
    return True
