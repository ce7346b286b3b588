"""CPU oracle for the ppvector speaker-embedding hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy / torch fp32+fp64), the arithmetic of the
reference hot path  waveform -> Kaldi Fbank-80 + CMN (or an STFT front end) -> ECAPA-TDNN /
ResNetSE / ERes2Net / CAM++ -> pooling -> 192-d embedding -> {AAMLoss | cosine scoring}, and of the
training step around it (train-mode BatchNorm, autograd backward, Adam).  It exists so that the CUDA path can be
checked against something; it is NOT part of the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import it;
  * the product (``voiceprintrecognition-paddlepaddle_b200/ppvector``) never imports
    it and fails loudly when the CUDA library is missing.

PARITY PIN: the reference ships no tests, golden vectors or weights and
PaddlePaddle is not installable here, so the pin is made in two parts:

  * MODEL / LOSS / SCHEDULE GRAPH -- pinned to the reference's own code.  ``tests/golden/make_ref_fixtures.py``
    imports ``/root/reference/ppvector/models/{utils,pooling,ecapa_tdnn,resnet_se,eres2net,campplus,fc}.py``,
    ``loss/aamloss.py`` and ``optimizer/scheduler.py`` UNMODIFIED under ``tests/paddle_shim`` (a paddle -> torch
    stand-in for the ~50 Paddle calls those files make), loads this oracle's seeded weights by ``state_dict`` name and
    records fp64 outputs (embeddings + layer taps for four backbones at T in {98, 298}, ``lengths`` / no-global-context /
    shortcut-conv variants, pooling modules, classifier + AAMLoss forward / backward, one TRAIN-mode step with
    gradients and running statistics, both schedules).  ``tests/test_oracle_vs_reference.py`` asserts this oracle equals
    those outputs to 1e-10.  What remains ASSUMED is only the semantics of each Paddle op, listed one by one in
    ``tests/paddle_shim/README.md``.
  * FRONT ENDS -- the arithmetic lives in un-vendored packages (paddleaudio>=1.0.1, paddle.audio, yeaudio>=0.0.6), so
    the pin is ``torchaudio.compliance.kaldi.fbank`` (torchaudio 2.11 -- the code paddleaudio's ``compliance/kaldi.py``
    was ported from) on the five wav files bundled with the reference and on seeded synthetic audio
    (``tests/golden/fbank_*.npz``), torchaudio's mel / DCT / ``torch.stft`` for the STFT front ends, and the reference's
    own ``metric/metrics.py`` for EER / minDCF.  SpecAugment / dB-normalisation follow RECALLED yeaudio behaviour:
    that part is "parity unpinned" and says so where it is used.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
