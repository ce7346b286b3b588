"""Pre-compute the front-end features of the train / enrolment / trials lists on the GPU and write them as .npy files next to
new list files (counterpart of the reference's extract_features.py; same options)."""
from cli_common import parse_options

OPTIONS = [
    ('configs', str, 'configs/ecapa_tdnn.yml', 'model / data configuration (YAML)'),
    ('save_dir', str, 'dataset/features', 'where the .npy feature files go'),
    ('max_duration', int, 100, 'longest audio, in seconds, that is featurised without cropping'),
]


def main(opt):
    from ppvector.trainer import PPVectorTrainer
    PPVectorTrainer(configs=opt.configs).extract_features(save_dir=opt.save_dir, max_duration=opt.max_duration)


if __name__ == '__main__':
    main(parse_options(__doc__, OPTIONS))
