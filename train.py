"""Drop-in for the reference's train.py (same arguments).  One process per GPU: start it with
``python -m torch.distributed.run --nproc-per-node N train.py ...`` for data-parallel training (the reference uses
``python -m paddle.distributed.launch``); the gradient all-reduce runs over NCCL."""
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
add_arg('data_augment_configs', str, None, '数据增强配置文件 (only spec_aug is implemented on the B200 path)')
add_arg("use_gpu", bool, True, '是否使用GPU训练')
add_arg("do_eval", bool, True, '训练时是否评估模型')
add_arg('save_model_path', str, 'models/', '模型保存的路径')
add_arg('log_dir', str, 'log/', '保存VisualDL日志文件的路径 (unused: VisualDL logging is out of scope)')
add_arg('resume_model', str, None, '恢复训练，当为None则不使用预训练模型')
add_arg('pretrained_model', str, None, '预训练模型的路径，当为None则不使用预训练模型')

if __name__ == '__main__':
    args = parser.parse_args()
    print_arguments(args=args)
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    trainer = PPVectorTrainer(configs=args.configs, use_gpu=args.use_gpu, data_augment_configs=args.data_augment_configs)
    trainer.train(save_model_path=args.save_model_path, log_dir=args.log_dir, resume_model=args.resume_model,
                  pretrained_model=args.pretrained_model, do_eval=args.do_eval)
