"""Cosine similarity of the speaker embeddings of two audio files, and the same / different decision at a threshold
(counterpart of the reference's infer_contrast.py; same options)."""
from cli_common import parse_options

OPTIONS = [
    ('configs', str, 'configs/ecapa_tdnn.yml', 'model / data configuration (YAML)'),
    ('use_gpu', bool, True, 'must stay True: this build has no CPU path'),
    ('audio_path1', str, 'dataset/a_1.wav', 'first utterance'),
    ('audio_path2', str, 'dataset/b_2.wav', 'second utterance'),
    ('threshold', float, 0.6, 'scores above this count as the same speaker'),
    ('model_path', str, 'models/EcapaTdnn_Fbank/best_model/', 'directory or file holding the weights'),
]


def main(opt):
    from ppvector.predict import PPVectorPredictor
    predictor = PPVectorPredictor(configs=opt.configs, model_path=opt.model_path, use_gpu=opt.use_gpu)
    score = predictor.contrast(opt.audio_path1, opt.audio_path2)
    verdict = 'the same speaker' if score > opt.threshold else 'different speakers'
    print(f'{opt.audio_path1} and {opt.audio_path2}: {verdict} (similarity {score:.5f}, threshold {opt.threshold})')


if __name__ == '__main__':
    main(parse_options(__doc__, OPTIONS))
