"""CPU: host-side glue in front of and behind the hot path -- wav decoding / level normalisation / resampling
(ppvector/data_utils/audio.py, the slice of yeaudio the reference's predict.py:189-216 and reader.py:85-104 use), checkpoint files
(ppvector/utils/checkpoint.py: the reference saves a pickled {name: ndarray} dict, `model.pdparams`), zero-padding collate
(collate_fn.py:5-23) and the command-line option tables."""
import io
import os
import pickle
import wave

import numpy as np
import pytest
import torch

from oracle import fbank as ofb
from ppvector.data_utils.audio import AudioSegment, normalize_db, read_wav, resample
from ppvector.data_utils.collate_fn import collate_fn
from ppvector.utils.checkpoint import load_state_dict_file


def _wav_bytes(x, sr=16000, width=2, channels=1):
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(width)
        w.setframerate(sr)
        if width == 2:
            w.writeframes((np.clip(x, -1, 1) * 32767).astype("<i2").tobytes())
        else:
            w.writeframes(((np.clip(x, -1, 1) * 127) + 128).astype(np.uint8).tobytes())
    return buf.getvalue()


def test_wav_decode_paths(tmp_path):
    g = np.random.RandomState(0)
    x = (0.3 * g.randn(4000)).astype(np.float32)
    raw = _wav_bytes(x)
    y, sr = read_wav(raw)
    assert sr == 16000 and y.dtype == np.float32 and np.abs(y - np.clip(x, -1, 1)).max() < 1e-4
    p = tmp_path / "a.wav"
    p.write_bytes(raw)
    assert np.array_equal(AudioSegment.from_file(str(p)).samples, y)
    assert np.array_equal(AudioSegment.from_bytes(raw).samples, y)
    stereo = np.stack([x, -x], axis=1).reshape(-1)
    ys, _ = read_wav(_wav_bytes(stereo, channels=2))
    assert np.abs(ys).max() < 1e-4  # channels are averaged
    y8, _ = read_wav(_wav_bytes(x, width=1))
    assert np.abs(y8 - np.clip(x, -1, 1)).max() < 2.0 / 127


def test_db_normalisation_matches_the_oracle_and_hits_the_target():
    g = np.random.RandomState(1)
    x = (0.05 * g.randn(16000)).astype(np.float32)
    y = normalize_db(x, -20.0)
    assert abs(10 * np.log10(np.mean(y.astype(np.float64) ** 2)) + 20.0) < 1e-4
    assert np.abs(y - ofb.db_normalize(x, -20.0)).max() < 1e-6
    with pytest.raises(ValueError):
        normalize_db(np.zeros(100, np.float32), -20.0)
    seg = AudioSegment.from_ndarray(x, 16000)
    seg.normalize(target_db=-20)
    assert np.array_equal(seg.samples, y) and abs(seg.duration - 1.0) < 1e-9


def test_resample_keeps_a_tone():
    t = np.arange(8000) / 8000.0
    x = np.sin(2 * np.pi * 440 * t).astype(np.float32)
    y = resample(x, 8000, 16000)
    assert abs(len(y) - 16000) <= 1
    t2 = np.arange(len(y)) / 16000.0
    assert np.abs(y[200:-200] - np.sin(2 * np.pi * 440 * t2)[200:-200]).max() < 2e-2
    assert resample(x, 16000, 16000) is x


def test_checkpoint_readers(tmp_path):
    sd = {"0.blocks.0.conv.conv.weight": np.arange(24, dtype=np.float32).reshape(2, 3, 4), "0.fc.conv.bias": np.ones(5, np.float32),
          "1.weight": np.zeros((5, 7), np.float32)}
    d = tmp_path / "best_model"
    d.mkdir()
    with open(d / "model.pdparams", "wb") as f:  # paddle.save of a state_dict: a pickle of {name: ndarray}
        pickle.dump(sd, f, protocol=2)
    got = load_state_dict_file(str(d))
    assert set(got) == set(sd) and all(np.array_equal(got[k], sd[k]) for k in sd)
    with open(tmp_path / "tuple.pdparams", "wb") as f:  # some paddle versions store (name, ndarray) pairs
        pickle.dump({k: (k, v) for k, v in sd.items()}, f, protocol=2)
    got = load_state_dict_file(str(tmp_path / "tuple.pdparams"))
    assert all(np.array_equal(got[k], sd[k]) for k in sd)
    np.savez(tmp_path / "model.npz", **sd)
    assert np.array_equal(load_state_dict_file(str(tmp_path / "model.npz"))["1.weight"], sd["1.weight"])
    torch.save({k: torch.from_numpy(v) for k, v in sd.items()}, tmp_path / "model.pt")
    assert np.array_equal(load_state_dict_file(str(tmp_path / "model.pt"))["0.fc.conv.bias"], sd["0.fc.conv.bias"])
    empty = tmp_path / "empty_model_dir"
    empty.mkdir()
    with pytest.raises(FileNotFoundError):
        load_state_dict_file(str(empty))


def test_collate_pads_to_the_longest():
    a, b = torch.ones(5, 3), 2 * torch.ones(8, 3)
    feats, labels, lens = collate_fn([(a, 4), (b, 9)])
    assert feats.shape == (2, 8, 3) and labels.tolist() == [4, 9] and lens.tolist() == [5, 8]
    assert torch.equal(feats[0, :5], a) and float(feats[0, 5:].abs().sum()) == 0.0 and torch.equal(feats[1], b)


def test_cli_option_tables(capsys):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import cli_common
    import importlib
    opt = cli_common.parse_options("x", [("configs", str, "c.yml", "h"), ("use_gpu", bool, True, "h"), ("n", int, 3, "h"), ("m", str, None, "h")],
                                   ["--use_gpu", "False", "--n", "7", "--m", "None"])
    assert opt.configs == "c.yml" and opt.use_gpu is False and opt.n == 7 and opt.m is None
    with pytest.raises(SystemExit):
        cli_common.parse_options("x", [("use_gpu", bool, True, "h")], ["--use_gpu", "maybe"])
    # the option names of each entry script are the reference's
    expect = {"train": {"configs", "data_augment_configs", "use_gpu", "do_eval", "save_model_path", "log_dir", "resume_model", "pretrained_model"},
              "eval": {"configs", "use_gpu", "save_image_path", "resume_model"},
              "extract_features": {"configs", "save_dir", "max_duration"},
              "infer_contrast": {"configs", "use_gpu", "audio_path1", "audio_path2", "threshold", "model_path"}}
    for name, keys in expect.items():
        mod = importlib.import_module(name)
        assert {row[0] for row in mod.OPTIONS} == keys, name


def test_accuracy_helpers():
    from ppvector.utils.utils import cal_accuracy, cal_accuracy_threshold, cosin_metric
    g = np.random.RandomState(3)
    truth = g.rand(500) < 0.4
    scores = np.where(truth, 0.62, 0.35) + 0.15 * g.randn(500)
    # the reference's loop, restated
    best_acc, best_thr = 0, 0
    for i in range(100):
        thr = i * 0.01
        acc = np.mean(((scores >= thr) == truth).astype(int))
        if acc > best_acc:
            best_acc, best_thr = acc, thr
    got_acc, got_thr = cal_accuracy_threshold(scores, truth.astype(int))
    assert abs(got_acc - best_acc) < 1e-12 and abs(got_thr - best_thr) < 1e-12
    assert abs(cal_accuracy(scores, truth, 0.5) - np.mean((scores >= 0.5) == truth)) < 1e-12
    a, b = g.randn(192), g.randn(192)
    assert abs(cosin_metric(a, b) - float(a @ b / np.linalg.norm(a) / np.linalg.norm(b))) < 1e-12


def _riff_bytes(samples_f64, sr, tag, bits, ch=1, extensible=False):
    """Build a RIFF/WAVE file in memory (test helper): tag 1 = PCM, 3 = IEEE float."""
    import struct
    x = np.asarray(samples_f64, dtype=np.float64).reshape(-1, ch)
    if tag == 3:
        raw = x.astype("<f4" if bits == 32 else "<f8").tobytes()
    elif bits == 8:
        raw = (np.round(x * 128.0) + 128).clip(0, 255).astype(np.uint8).tobytes()
    elif bits == 16:
        raw = np.round(x * 32768.0).clip(-32768, 32767).astype("<i2").tobytes()
    elif bits == 24:
        v = np.round(x * 8388608.0).clip(-8388608, 8388607).astype(np.int64).reshape(-1)
        v = np.where(v < 0, v + (1 << 24), v)
        raw = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255], axis=1).astype(np.uint8).tobytes()
    else:
        raw = np.round(x * 2147483648.0).clip(-2147483648, 2147483647).astype("<i4").tobytes()
    block = ch * bits // 8
    if extensible:
        guid_tail = bytes.fromhex("000000001000800000aa00389b71")
        fmt = struct.pack("<HHIIHH", 0xFFFE, ch, sr, sr * block, block, bits) + struct.pack("<HHI", 22, bits, 0) + struct.pack("<H", tag) + guid_tail
    else:
        fmt = struct.pack("<HHIIHH", tag, ch, sr, sr * block, block, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"data" + struct.pack("<I", len(raw)) + raw
    return b"RIFF" + struct.pack("<I", len(body)) + body


@pytest.mark.parametrize("tag,bits,tol", [(1, 8, 1 / 100), (1, 16, 1e-4), (1, 24, 1e-6), (1, 32, 1e-7), (3, 32, 1e-7), (3, 64, 1e-7)])
def test_wav_decoder_formats(tag, bits, tol):
    """AudioSegment.from_file / from_bytes read what yeaudio reads through soundfile for WAV: integer PCM 8-32 bit (incl. packed
    24-bit), IEEE float, WAVE_FORMAT_EXTENSIBLE headers, extra chunks, multi-channel (averaged); other containers raise by name."""
    from ppvector.data_utils.audio import AudioSegment, read_wav
    rng = np.random.default_rng(bits + tag)
    x = rng.uniform(-0.9, 0.9, 1000)
    for ext in (False, True):
        y, sr = read_wav(_riff_bytes(x, 22050, tag, bits, extensible=ext))
        assert sr == 22050 and y.dtype == np.float32 and y.shape == (1000,)
        assert np.abs(y - x).max() < tol
    st = rng.uniform(-0.5, 0.5, (500, 2))
    y, _ = read_wav(_riff_bytes(st, 16000, tag, bits, ch=2))
    assert np.abs(y - st.mean(1)).max() < 2 * tol
    seg = AudioSegment.from_bytes(_riff_bytes(x, 16000, tag, bits))
    assert seg.sample_rate == 16000 and abs(seg.duration - 1000 / 16000) < 1e-9


def test_wav_decoder_rejects_other_containers_by_name(tmp_path):
    from ppvector.data_utils.audio import read_wav
    try:
        import soundfile  # noqa: F401
        pytest.skip("soundfile installed: other containers are decoded by it")
    except ImportError:
        pass
    with pytest.raises(ValueError, match="FLAC"):
        read_wav(b"fLaC" + bytes(64))
    with pytest.raises(ValueError, match="format tag"):
        import struct
        fmt = struct.pack("<HHIIHH", 0x0055, 1, 16000, 2000, 1, 0)
        body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"data" + struct.pack("<I", 4) + bytes(4)
        read_wav(b"RIFF" + struct.pack("<I", len(body)) + body)


def _reference_windows(n, chunk_len, chunk_shift):
    """the windowing rule of the reference written as its sequential loop (infer_utils/speaker_diarization.py:62-86): the spec for chunk_windows"""
    out, last = [], 0
    for st in range(0, n, chunk_shift):
        ed = min(st + chunk_len, n)
        if ed <= last:
            break
        last = ed
        out.append((max(0, ed - chunk_len), ed))
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


def test_diarization_chunk_fan_out():
    """1.5 s windows every 0.75 s over the voiced segments, the last window moved back to end at the segment end, short
    segments zero-padded; every window goes to the embedding path as one equal-length batch (ppvector/infer_utils/chunking.py)."""
    from ppvector.infer_utils.chunking import chunk_segments, chunk_windows, fan_out_embeddings
    rng = np.random.default_rng(3)
    for n in [0, 1, 11999, 12000, 23999, 24000, 24001, 36000, 64000, 64001, 100000] + list(rng.integers(1, 200000, size=40)):
        assert np.array_equal(chunk_windows(int(n), 24000, 12000), _reference_windows(int(n), 24000, 12000)), n
    assert np.array_equal(chunk_windows(50, 7, 3), _reference_windows(50, 7, 3))
    # a 4.0 s segment starting at 2.0 s: four windows, the last one moved back to end at 6.0 s
    x = rng.standard_normal(64000).astype(np.float32)
    short = rng.standard_normal(8000).astype(np.float32)
    times, chunks = chunk_segments([(2.0, 6.0, x), (7.5, 8.0, short)])
    assert np.allclose(times, [[2.0, 3.5], [2.75, 4.25], [3.5, 5.0], [4.25, 5.75], [4.5, 6.0], [7.5, 8.0]])
    assert chunks.shape == (6, 24000) and chunks.dtype == np.float32
    assert np.array_equal(chunks[1], x[12000:36000]) and np.array_equal(chunks[4], x[40000:64000])
    assert np.array_equal(chunks[5][:8000], short) and not chunks[5][8000:].any()
    # fan-out: the embedding function sees equal-length batches of at most batch_size rows, results come back in window order
    seen = []

    def embed(b):
        seen.append(b.shape)
        return np.stack([b.mean(1), b.std(1)], axis=1)

    t2, e = fan_out_embeddings([(2.0, 6.0, x), (7.5, 8.0, short)], embed, batch_size=4)
    assert seen == [(4, 24000), (2, 24000)] and np.array_equal(t2, times)
    assert np.allclose(e, np.stack([chunks.mean(1), chunks.std(1)], axis=1))
    t0, e0 = fan_out_embeddings([], embed)
    assert t0.shape == (0, 2) and e0.shape[0] == 0


def test_diarization_embeddings_wrapper_plumbing():
    """PPVectorPredictor.diarization_embeddings = _load_audio (dB normalisation etc.) + the chunk fan-out + extract_embeddings.  The embedding call
    needs the GPU; here it is stubbed to check what the wrapper hands to it (shapes, order, VAD spans, batch splitting)."""
    import os

    import yaml

    from ppvector.predict import PPVectorPredictor
    from ppvector.utils.utils import dict_to_object
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = PPVectorPredictor.__new__(PPVectorPredictor)  # no CUDA in this test: skip __init__
    p.configs = dict_to_object(yaml.load(open(os.path.join(root, "configs", "ecapa_tdnn.yml")), Loader=yaml.FullLoader))
    calls = []

    def stub(w, ratio=None):
        calls.append(w.shape)
        return np.stack([w.mean(1), w.std(1)], 1).astype(np.float32)

    p.extract_embeddings = stub
    x = (0.1 * np.random.default_rng(5).standard_normal(80000)).astype(np.float32)
    t, e = p.diarization_embeddings(x, sample_rate=16000, batch_size=4)
    assert np.allclose(t, [[0, 1.5], [0.75, 2.25], [1.5, 3.0], [2.25, 3.75], [3.0, 4.5], [3.5, 5.0]]) and e.shape == (6, 2)
    assert calls == [(4, 24000), (2, 24000)]
    t2, e2 = p.diarization_embeddings(x, sample_rate=16000, vad_segments=[(0.5, 2.6), (3.0, 3.4)])
    assert np.allclose(t2, [[0.5, 2.0], [1.1, 2.6], [3.0, 3.4]]) and e2.shape == (3, 2)
