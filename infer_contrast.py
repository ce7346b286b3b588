"""Drop-in for the reference's infer_contrast.py (same arguments): cosine similarity of two utterances."""
import argparse
import functools
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'voiceprintrecognition-paddlepaddle_b200'))
from ppvector.predict import PPVectorPredictor  # noqa: E402
from ppvector.utils.utils import add_arguments, print_arguments  # noqa: E402

parser = argparse.ArgumentParser(description=__doc__)
add_arg = functools.partial(add_arguments, argparser=parser)
add_arg('configs', str, 'configs/ecapa_tdnn.yml', '配置文件')
add_arg('use_gpu', bool, True, '是否使用GPU预测')
add_arg('audio_path1', str, 'dataset/a_1.wav', '预测第一个音频')
add_arg('audio_path2', str, 'dataset/b_2.wav', '预测第二个音频')
add_arg('threshold', float, 0.6, '判断是否为同一个人的阈值')
add_arg('model_path', str, 'models/EcapaTdnn_Fbank/best_model/', '导出的预测模型文件路径')

if __name__ == '__main__':
    args = parser.parse_args()
    print_arguments(args=args)
    predictor = PPVectorPredictor(configs=args.configs, model_path=args.model_path, use_gpu=args.use_gpu)
    dist = predictor.contrast(args.audio_path1, args.audio_path2)
    if dist > args.threshold:
        print(f"{args.audio_path1} 和 {args.audio_path2} 为同一个人，相似度为：{dist}")
    else:
        print(f"{args.audio_path1} 和 {args.audio_path2} 不是同一个人，相似度为：{dist}")
