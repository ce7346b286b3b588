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

PARITY UNPINNED: the reference ships no tests, no golden
vectors and no weights, and its arithmetic lives in un-vendored
third-party packages that are not installable here (paddlepaddle 2.5/2.6,
paddleaudio>=1.0.1, yeaudio>=0.0.6).  What the oracle *is* pinned against:

  * ``torchaudio.compliance.kaldi.fbank`` (torchaudio 2.11, present in this image) --
    the code paddleaudio's ``compliance/kaldi.py`` was ported from -- on the five
    wav files bundled with the reference and on seeded synthetic audio
    (``tests/golden/fbank_*.npz``, made by ``tests/golden/make_golden.py``);
  * ``torch.nn.functional`` fp64 evaluation of the same graph for the model;
  * the structural known-answers the reference README prints (``paddle.summary``
    parameter counts, README.md:303-351).

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
