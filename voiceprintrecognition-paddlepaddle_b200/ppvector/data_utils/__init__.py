"""Audio in, features out: wav decoding and level normalisation (audio.py), list-file datasets (reader.py), batching (collate_fn.py),
GPU front ends (featurizer.py) and SpecAugment masking (spec_aug.py)."""
