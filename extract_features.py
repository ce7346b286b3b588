"""Drop-in for the reference's extract_features.py (same arguments): Fbank features of the train / enrol / trials lists
to .npy files, computed on the GPU."""
import argparse
import functools
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'voiceprintrecognition-paddlepaddle_b200'))
from ppvector.trainer import PPVectorTrainer  # noqa: E402
from ppvector.utils.utils import add_arguments, print_arguments  # noqa: E402

parser = argparse.ArgumentParser(description=__doc__)
add_arg = functools.partial(add_arguments, argparser=parser)
add_arg('configs', str, 'configs/ecapa_tdnn.yml', '配置文件')
add_arg('save_dir', str, 'dataset/features', '保存特征的路径')
add_arg('max_duration', int, 100, '提取特征的最大时长，单位秒')

if __name__ == '__main__':
    args = parser.parse_args()
    print_arguments(args=args)
    trainer = PPVectorTrainer(configs=args.configs)
    trainer.extract_features(save_dir=args.save_dir, max_duration=args.max_duration)
