"""Scoring and verification metrics: cosine scoring on the GPU (cosine.py), EER / minDCF on the host (metrics.py)."""
