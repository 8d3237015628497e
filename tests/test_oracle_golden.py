"""Pins the oracle (and the host-side mirror modules) to every golden vector the reference's
own tests hold for this path, plus TensorFlow's own CTC known-answer vectors."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import keras_tf_oracle as oracle
from speechless_b200 import english_frequent_characters, german_frequent_characters
from speechless_b200.grapheme_enconding import AsgGraphemeEncoding, CtcGraphemeEncoding

GOLDEN = Path(__file__).parent / "golden"


# ---- reference speechless/test/test_ctc_decoders.py:22-24,38-41 ("A A blank A A", V=2, blank=1)
def test_greedy_decoder_reference_vector():
    logits = np.array([[[1.0, 0.0]], [[1.0, 0.0]], [[0.0, 1.0]], [[1.0, 0.0]], [[1.0, 0.0]]], dtype=np.float32)
    probs = logits.transpose(1, 0, 2)  # time-major in the reference test -> (B, T, V)
    merged, merged_len = oracle.greedy_decode(probs, [5], blank=1, merge_repeated=True)
    unmerged, unmerged_len = oracle.greedy_decode(probs, [5], blank=1, merge_repeated=False)
    assert list(merged[0, :merged_len[0]]) == [0, 0]
    assert list(unmerged[0, :unmerged_len[0]]) == [0, 0, 0, 0]
    assert (merged[0, merged_len[0]:] == -1).all()


# ---- reference speechless/test/test_grapheme_encoding.py:9-31
def test_ctc_grapheme_encoding_reference_vectors():
    g = CtcGraphemeEncoding(english_frequent_characters)
    assert g.grapheme_set_size == 29 and g.ctc_blank == 28
    label = "she wasn't three abcxyz"
    assert g.decode_graphemes(g.encode(label), merge_repeated=False) == label
    graphemes = g.encode("sssshhhheeeee      wasn't thre") + [g.ctc_blank] + g.encode("eeeeee")
    assert g.decode_graphemes(graphemes) == "she wasn't three"
    predictions = np.zeros((2, 3, g.grapheme_set_size))
    for row in range(2):
        for t, c in enumerate("abc"):
            predictions[row, t, g.encode_character(c)] = 1
    assert g.decode_prediction_batch(predictions, prediction_lengths=[3, 2]) == ["abc", "ab"]
    # the oracle's own decode agrees
    assert oracle.decode_graphemes(graphemes, english_frequent_characters) == "she wasn't three"


# ---- reference speechless/test/test_grapheme_encoding.py:34-50
def test_asg_grapheme_encoding_reference_vectors():
    g = AsgGraphemeEncoding(english_frequent_characters)
    assert g.encode("ee") == [g.encode_character("e"), g.asg_twice]
    assert g.encode("eee") == [g.encode_character("e"), g.asg_thrice]
    with pytest.raises(ValueError):
        g.encode("eeee")
    chars = lambda s: [g.encode_character(c) for c in s]
    graphemes = chars("sssshhhheeeee      wasn't thre") + [g.asg_twice] * 3 + chars("    aaaaaaa") + [g.asg_thrice]
    assert g.decode_graphemes(graphemes) == "she wasn't three aaa"


def test_grapheme_errors_and_label_batch():
    g = CtcGraphemeEncoding(german_frequent_characters)
    assert g.grapheme_set_size == 33
    with pytest.raises(ValueError, match="Unexpected char"):
        g.encode("Ä")
    with pytest.raises(ValueError, match="Unexpected grapheme"):
        g.decode_graphemes([40])
    batch = g.encode_label_batch(["ab", "ößa", "z"])
    assert batch.dtype == np.int32 and batch.shape == (3, 3)
    assert batch.tolist() == [[0, 1, -1], [29, 31, 0], [25, -1, -1]]
    assert (oracle.encode_label_batch(["ab", "ößa", "z"], german_frequent_characters) == batch).all()


def test_golden_fixture_from_reference_module():
    """tests/golden/grapheme_reference.json was produced by importing the reference's own
    speechless/grapheme_enconding.py (tests/golden/make_grapheme_golden.py)."""
    data = json.loads((GOLDEN / "grapheme_reference.json").read_text())
    for case in data["cases"]:
        alphabet = list(case["alphabet"])
        g = CtcGraphemeEncoding(alphabet)
        assert g.encode_label_batch(case["labels"]).tolist() == case["label_batch"]
        assert [g.encode(l) for l in case["labels"]] == case["encoded"]
        for graphemes, merged, unmerged in case["decodes"]:
            assert g.decode_graphemes(graphemes) == merged
            assert g.decode_graphemes(graphemes, merge_repeated=False) == unmerged
            assert oracle.decode_graphemes(graphemes, alphabet) == merged
        scores = np.array(case["prediction_batch"])
        assert g.decode_prediction_batch(scores, case["prediction_lengths"]) == case["decoded_predictions"]
        dense, lens = oracle.greedy_decode(scores, case["prediction_lengths"])
        ours = [oracle.decode_graphemes(list(dense[i, :lens[i]]), alphabet, merge_repeated=False)
                for i in range(len(lens))]
        assert ours == case["decoded_predictions"]


# ---- TensorFlow's ctc_loss_op_test.py testBasic known answers (the third-party op behind
#      K.ctc_batch_cost, net.py:402-406): 6 classes, blank = 5
TF_PROBS_0 = np.array([[0.633766, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
                       [0.111121, 0.588392, 0.278779, 0.0055756, 0.00569609, 0.010436],
                       [0.0357786, 0.633813, 0.321418, 0.00249248, 0.00272882, 0.0037688],
                       [0.0663296, 0.643849, 0.280111, 0.00283995, 0.0035545, 0.00331533],
                       [0.458235, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]])
TF_PROBS_1 = np.array([[0.30176, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
                       [0.24082, 0.397533, 0.0557226, 0.0546814, 0.0557528, 0.19549],
                       [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, 0.202456],
                       [0.280884, 0.429522, 0.0326593, 0.0339046, 0.0326856, 0.190345],
                       [0.423286, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]])


@pytest.mark.parametrize("probs,label,neg_log_prob", [(TF_PROBS_0, [0, 1, 2, 1, 0], 3.34211),
                                                      (TF_PROBS_1, [0, 1, 1, 0], 5.42262)])
def test_ctc_loss_tensorflow_known_answers(probs, label, neg_log_prob):
    lp = np.log(probs)
    _, _, ll, _ = oracle.ctc_alpha_beta(lp, label, blank=5)
    assert abs(-ll - neg_log_prob) < 2e-5
    assert abs(oracle.ctc_brute_force_log_likelihood(lp, label, blank=5) - ll) < 1e-9


def test_ctc_gradient_tensorflow_known_answer():
    """testBasic's first example forces one alignment (5 labels in 5 frames), so TF's expected
    gradient wrt the unnormalised inputs is softmax - onehot(label)."""
    label = [0, 1, 2, 1, 0]
    lp = np.log(TF_PROBS_0)
    alpha, beta, ll, ext = oracle.ctc_alpha_beta(lp, label, blank=5)
    occ = np.zeros_like(lp)
    for s, v in enumerate(ext):
        occ[:, v] += np.exp(alpha[:, s] + beta[:, s] - lp[:, v] - ll)
    grad = np.exp(lp) - occ
    expected = TF_PROBS_0.copy()
    expected[np.arange(5), label] -= 1.0
    assert np.abs(grad - expected).max() < 1e-5
    assert abs(expected[0, 0] - (-0.366234)) < 1e-6 and abs(expected[4, 0] - (-0.541765)) < 1e-6


# ---- TF SAME padding values checked in SURVEY.md A.1
@pytest.mark.parametrize("T,k,s,expected", [(1251, 48, 2, (626, 23, 24)), (1250, 48, 2, (625, 23, 23)),
                                            (626, 7, 1, (626, 3, 3)), (626, 32, 1, (626, 15, 16)),
                                            (626, 1, 1, (626, 0, 0)), (160000, 250, 160, (1000, 45, 45))])
def test_same_padding(T, k, s, expected):
    assert oracle.same_padding(T, k, s) == expected


def test_layer_specs_match_reference_architecture():
    specs = oracle.wav2letter_layer_specs(128, 29)
    assert [s[0] for s in specs] == ["striding_conv"] + ["inner_conv_%d" % i for i in range(1, 8)] + [
        "big_conv_1", "big_conv_2", "output_conv"]
    params = sum(k * cin * cout + cout for (_, cin, cout, k, _, _) in specs)
    assert params == 24662529
    from speechless_b200.engine import wav2letter_layers
    ours = wav2letter_layers(128, 29)
    assert [(l.name, l.cin, l.cout, l.kernel, l.stride, l.activation) for l in ours] == specs
