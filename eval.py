"""Drop-in for the reference's eval.py (same arguments): EER / minDCF of a checkpoint on the enrol / trials lists."""
import argparse
import functools
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'voiceprintrecognition-paddlepaddle_b200'))
from ppvector.trainer import PPVectorTrainer  # noqa: E402
from ppvector.utils.utils import add_arguments, print_arguments  # noqa: E402

parser = argparse.ArgumentParser(description=__doc__)
add_arg = functools.partial(add_arguments, argparser=parser)
add_arg('configs', str, 'configs/ecapa_tdnn.yml', "配置文件")
add_arg("use_gpu", bool, True, "是否使用GPU评估模型")
add_arg('save_image_path', str, 'output/images/', "保存结果图的路径")
add_arg('resume_model', str, 'models/EcapaTdnn_Fbank/best_model/', "模型的路径")

if __name__ == '__main__':
    args = parser.parse_args()
    print_arguments(args=args)
    trainer = PPVectorTrainer(configs=args.configs, use_gpu=args.use_gpu)
    start = time.time()
    eer, min_dcf, threshold = trainer.evaluate(resume_model=args.resume_model, save_image_path=args.save_image_path)
    print('评估消耗时间：{}s，threshold：{:.2f}，EER: {:.5f}, MinDCF: {:.5f}'.format(int(time.time() - start), threshold, eer, min_dcf))
