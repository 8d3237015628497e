"""Host-side pieces that need no GPU: result types, tools, the C-ABI library's symbol table."""
import ctypes
import json
import math
import re
from pathlib import Path

import numpy as np
import pytest

from speechless_b200 import _lib
from speechless_b200.results import (ExpectationVsPrediction, ExpectationsVsPredictions,
                                     ExpectationsVsPredictionsInBatches, ExpectationsVsPredictionsInGroupedBatches,
                                     levenshtein)
from speechless_b200.tools import average_or_nan, paginate, single, single_or_none

ROOT = Path(__file__).resolve().parent.parent


def test_paginate_reference_vector():
    # reference speechless/test/test_spectrogram_batch.py:6-9
    assert list(paginate([1, 2, 3], 2)) == [[1, 2], [3]]


def test_tools():
    assert single([7]) == 7
    with pytest.raises(AssertionError):
        single([1, 2])
    assert single_or_none([]) is None and single_or_none([3]) == 3
    with pytest.raises(AssertionError):
        single_or_none([1, 2])
    assert math.isnan(average_or_nan([])) and average_or_nan([1, 2, 6]) == 3


def test_levenshtein():
    assert levenshtein("kitten", "sitting") == 3
    assert levenshtein("", "abc") == 3 and levenshtein("abc", "") == 3 and levenshtein("abc", "abc") == 0
    assert levenshtein("a b c".split(), "a x c d".split()) == 2
    rng = np.random.default_rng(0)
    for _ in range(50):  # against the textbook full-matrix DP
        a = "".join(rng.choice(list("abc"), size=rng.integers(0, 9)))
        b = "".join(rng.choice(list("abc"), size=rng.integers(0, 9)))
        d = np.zeros((len(a) + 1, len(b) + 1), dtype=int)
        d[:, 0] = np.arange(len(a) + 1)
        d[0, :] = np.arange(len(b) + 1)
        for i in range(1, len(a) + 1):
            for j in range(1, len(b) + 1):
                d[i, j] = min(d[i - 1, j] + 1, d[i, j - 1] + 1, d[i - 1, j - 1] + (a[i - 1] != b[j - 1]))
        assert levenshtein(a, b) == d[-1, -1]


def test_result_types_reference_smoke_and_formats():
    # reference speechless/test/test_net.py:9-21 (smoke: nested groups incl. empty ones must print)
    a = ExpectationVsPrediction(expected="A", predicted="A", loss=0.0)
    b = ExpectationVsPrediction(expected="B", predicted="A", loss=2.0)
    batches = [ExpectationsVsPredictions([a, b]), ExpectationsVsPredictions([])]
    by_name = ExpectationsVsPredictionsInBatches(result_batches=batches)
    grouped = ExpectationsVsPredictionsInGroupedBatches(results_by_group_name=dict([
        ("corpus1", by_name), ("corpus2", by_name), ("empty", ExpectationsVsPredictionsInBatches([]))]))
    text = str(grouped)
    assert "corpus1: All batches: Average over 2 examples: 0.5 letter errors (50.00%), 0.5 word errors (50.00%), loss 1.00." in text
    assert "empty: All batches: Average over 0 examples: nan letter errors (nan%)" in text
    assert "All corpora: Average over 4 examples" in text
    # exact per-example format of reference net.py:47-52
    r = ExpectationVsPrediction(expected="the cat sat", predicted="the bat sat on", loss=37.188)
    assert str(r) == ('Expected:  "the cat sat"\nPredicted: "the bat sat on"\n'
                      'Errors: 4 letters (36%), 2 words (67%), loss: 37.19.')
    assert r.letter_error_count == 4 and r.word_error_count == 2


def test_header_symbols_are_exported_and_bound():
    """Every function include/speechless_b200.h declares is exported by the built library and
    has a ctypes signature (no compute calls here: there is no GPU)."""
    header = (ROOT / "include" / "speechless_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(sl_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    path = _lib.library_path()
    assert path.exists(), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(str(path))
    for name in declared:
        assert hasattr(lib, name), "{} missing from {}".format(name, path.name)
        assert name in _lib.SIGNATURES, "{} has no ctypes signature".format(name)
    assert sorted(_lib.SIGNATURES) == declared
    loaded = _lib.load()
    assert loaded.sl_version() >= 100
    # argument validation happens before any CUDA call
    assert loaded.sl_adam_step(None, None, None, None, 0, 1e-4, 0.9, 0.999, 1e-8, 1, None) == 1
    assert "null pointer" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(loaded.sl_ctc_greedy_decode(None, None, None, None, 1, 1, 2, 1, 1, None))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("SPEECHLESS_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_wav2letter_requires_cuda_device():
    import torch
    from speechless_b200 import english_frequent_characters
    from speechless_b200.net import Wav2Letter
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Wav2Letter(128, english_frequent_characters)
    with pytest.raises(ValueError, match="cannot be frozen"):
        Wav2Letter(128, english_frequent_characters, frozen_layer_count=3)
    assert Wav2Letter.model_file_name(12) == "weights-epoch12.h5"
    assert Wav2Letter.InputNames.prediction_lengths == "prediction_lenghts"  # typos are part of the surface
    assert Wav2Letter.InputNames.label_lengths == "label_lenghts"
    mapping = Wav2Letter.indices_to_load_by_target_index(list("abc"), list("cbx"))
    assert mapping == [2, 1, None]


def test_product_does_not_import_oracle():
    for path in (ROOT / "speechless_b200").rglob("*.py"):
        assert "oracle" not in path.read_text(), "{} must not reference the oracle".format(path)


def test_synthetic_shapes():
    from speechless_b200 import english_frequent_characters
    from speechless_b200.synthetic import frames_for_seconds, label_length_for, synthetic_batch
    assert frames_for_seconds(10) == 1251 and frames_for_seconds(60) == 7501
    assert label_length_for(1251) == 150 and label_length_for(7501) == 900
    batch = synthetic_batch(2, [30, 41], english_frequent_characters, seed=1)
    assert batch[1].z_normalized_transposed_spectrogram().shape == (41, 128)
    assert all(c in english_frequent_characters for c in batch[0].label)


def test_layer_table_with_raw_wave_input():
    """wave_conv k250 s160 in front (net.py:310-316); as a GEMM it is one tap over 250*Cin channels."""
    from speechless_b200.engine import same_padding, wav2letter_layers
    layers = wav2letter_layers(1, 29, use_raw_wave_input=True)
    assert [l.name for l in layers][:2] == ["wave_conv", "striding_conv"] and len(layers) == 12
    wave, striding = layers[0], layers[1]
    assert (wave.cin, wave.cout, wave.kernel, wave.stride, wave.windowed) == (1, 250, 250, 160, True)
    assert (wave.gemm_cin, wave.gemm_kernel, wave.gemm_stride, wave.cin_pad) == (250, 1, 1, 256)
    assert wave.w_size == 256 * 256 and striding.cin == 250 and striding.w_offset == wave.w_size + 256
    assert same_padding(160000, 250, 160) == (1000, 45)  # TF SAME: total 90 -> (45, 45)
    assert same_padding(160001, 250, 160) == (1001, 124)
    plain = wav2letter_layers(128, 29)
    assert len(plain) == 11 and not any(l.windowed for l in plain)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm) on a tiny sample: one JSON line on stdout with the
    contract's keys, the same `config` record the GPU arm prints, and `e2e` / `cpu_baseline` describing this run."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--workload", "small",
                          "--batch-per-gpu", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["steps"] == 1 and line["warmup"] == 0 and line["vs_baseline"] is None
    assert line["config"]["global_batch"] == 2 and line["config"]["frames_per_utterance"] == 1251
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["value"] > 0


def test_epoch_callback_writes_tensorboard_events_and_a_json_log(tmp_path):
    """The reference's TensorBoard callback (net.py:574): scalars per epoch, readable by TensorBoard itself."""
    from speechless_b200.net import Wav2Letter
    calls = []
    scalar_log, custom = Wav2Letter.create_callbacks(object(), lambda: calls.append(1), tmp_path / "tb", tmp_path / "nets",
                                                     callback_step=2, save=False)
    for epoch, loss in enumerate([3.5, 2.25, 1.125]):
        scalar_log(epoch, {"loss": loss})
        custom(epoch)
    assert len(calls) == 2  # epochs 0 and 2
    lines = [json.loads(line) for line in (tmp_path / "tb" / "scalars.jsonl").read_text().splitlines()]
    assert [(l["epoch"], l["loss"]) for l in lines] == [(0, 3.5), (1, 2.25), (2, 1.125)]
    pytest.importorskip("tensorboard")
    from tensorboard.backend.event_processing.event_file_loader import RawEventFileLoader
    from tensorboard.compat.proto.event_pb2 import Event
    event_files = sorted((tmp_path / "tb").glob("events.out.tfevents.*"))
    assert len(event_files) == 1
    events = [Event.FromString(raw) for raw in RawEventFileLoader(str(event_files[0])).Load()]
    scalars = [(e.step, v.tag, v.simple_value) for e in events for v in e.summary.value]
    assert scalars == [(0, "loss", 3.5), (1, "loss", 2.25), (2, "loss", 1.125)]
