"""Train a speaker-embedding model with the CUDA training step (counterpart of the reference's train.py; same options).
Data-parallel training: one process per GPU, ``python -m torch.distributed.run --nproc-per-node N train.py ...``; the single
gradient all-reduce per step runs over NCCL."""
from cli_common import init_distributed_if_launched, parse_options

OPTIONS = [
    ('configs', str, 'configs/cam++.yml', 'model / data configuration (YAML); the CUDA training step exists for configs/ecapa_tdnn.yml'),
    ('data_augment_configs', str, 'configs/augmentation.yml', 'augmentation configuration (YAML): speed / volume / noise / spec_aug on the GPU'),
    ('use_gpu', bool, True, 'must stay True: this build has no CPU path'),
    ('do_eval', bool, True, 'evaluate on the enrolment / trials lists after every epoch'),
    ('save_model_path', str, 'models/', 'where checkpoints are written'),
    ('log_dir', str, 'log/', 'accepted for compatibility; VisualDL logging is not produced'),
    ('resume_model', str, None, 'checkpoint to continue from'),
    ('pretrained_model', str, None, 'weights to start from'),
]


def main(opt):
    init_distributed_if_launched()
    from ppvector.trainer import PPVectorTrainer
    trainer = PPVectorTrainer(configs=opt.configs, use_gpu=opt.use_gpu, data_augment_configs=opt.data_augment_configs)
    trainer.train(save_model_path=opt.save_model_path, log_dir=opt.log_dir, resume_model=opt.resume_model,
                  pretrained_model=opt.pretrained_model, do_eval=opt.do_eval)


if __name__ == '__main__':
    main(parse_options(__doc__, OPTIONS))
