"""CTC prefix beam search: the CPU oracle against the reference's own vectors and brute force, the
word n-gram re-scorer, and (gpu) the CUDA kernel against the oracle through the C-ABI."""
import numpy as np
import pytest

from oracle import beam_search_oracle as bso

# reference speechless/test/test_ctc_decoders.py:22-24: "A A blank A A", V = 2, blank = 1
AA_BLANK_AA = np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 0.0]], dtype=np.float32)


def test_oracle_reproduces_the_reference_beam_search_rows():
    # test_ctc_decoders.py:38-39: beam_width=1 -> [0] with merge_repeated, [0, 0] without
    (merged, _), = bso.beam_search_decode(AA_BLANK_AA, beam_width=1, merge_repeated=True)
    (unmerged, _), = bso.beam_search_decode(AA_BLANK_AA, beam_width=1, merge_repeated=False)
    assert merged == [0]
    assert unmerged == [0, 0]


def test_wide_beam_is_exact_against_path_enumeration():
    rng = np.random.default_rng(0)
    for _ in range(40):
        T, V = int(rng.integers(1, 7)), int(rng.integers(2, 5))
        scores = rng.normal(size=(T, V)) * 2
        totals = bso.brute_force_labelings(scores)
        best = max(totals, key=totals.get)
        ranked = bso.beam_search_decode(scores, beam_width=10000, top_paths=3, merge_repeated=False)
        assert tuple(ranked[0][0]) == best
        assert ranked[0][1] == pytest.approx(totals[best], abs=1e-9)
        for labels, log_probability in ranked:  # every reported hypothesis carries its exact total
            assert log_probability == pytest.approx(totals[tuple(labels)], abs=1e-9)


def test_beam_hypothesis_scores_are_bounded_by_the_ctc_forward_likelihood():
    """Two oracles against each other: the score a beam search reports for a labeling is the mass of the
    alignments it kept, so it can never exceed log p(labeling | x) of the CTC forward algorithm
    (keras_tf_oracle.ctc_alpha_beta) and equals it when no relevant prefix was ever pruned."""
    from oracle import keras_tf_oracle as oracle
    rng = np.random.default_rng(11)
    equal = 0
    for _ in range(25):
        T, V = int(rng.integers(5, 40)), int(rng.integers(3, 8))
        scores = rng.normal(size=(T, V)) * rng.uniform(1, 4)
        lp = bso.log_softmax(scores)
        for width in (2, 8, 64):
            for tf_like in (True, False):
                for labels, log_probability in bso.beam_search_decode(scores, beam_width=width, top_paths=3,
                                                                     merge_repeated=False, tf_deactivation=tf_like):
                    exact = oracle.ctc_alpha_beta(lp, labels, V - 1)[2] if labels else float(lp[:, V - 1].sum())
                    assert log_probability <= exact + 1e-9
                    equal += abs(log_probability - exact) < 1e-9
    assert equal > 0  # (some hypotheses keep all of their alignments even in narrow beams)


def test_tf_deactivation_only_matters_for_narrow_beams():
    """The order-independent rule the kernel implements equals TF's for beam_width = 1 and for beams
    that never overflow; in between it may keep hypotheses TF drops (never a worse best path)."""
    rng = np.random.default_rng(3)
    differing = 0
    for _ in range(60):
        T, V, W = int(rng.integers(4, 30)), int(rng.integers(2, 7)), int(rng.integers(2, 7))
        scores = rng.normal(size=(T, V)) * rng.uniform(0.5, 4)
        for width in (1, W, 100000):
            if width == 100000:
                scores = scores[:6, :4]  # (every prefix stays in the beam: keep the tree small)
            tf_like = bso.beam_search_decode(scores, beam_width=width, merge_repeated=False)
            plain = bso.beam_search_decode(scores, beam_width=width, merge_repeated=False, tf_deactivation=False)
            if width == W:
                differing += tf_like[0][0] != plain[0][0]
                continue
            assert tf_like[0][0] == plain[0][0] and tf_like[0][1] == pytest.approx(plain[0][1], abs=1e-9)
    assert differing < 30  # (informative: a minority of the narrow-beam cases)


def test_beam_width_one_differs_from_greedy_as_the_reference_documents():
    # greedy (merge_repeated=True) gives [0, 0] on the same input: test_ctc_decoders.py:40
    from oracle import keras_tf_oracle as oracle
    probabilities = np.exp(bso.log_softmax(AA_BLANK_AA.astype(np.float64)))[None]
    dense, lengths = oracle.greedy_decode(probabilities, [5])
    assert dense[0, :lengths[0]].tolist() == [0, 0]


ARPA = """\\data\\
ngram 1=6
ngram 2=3

\\1-grams:
-1.0\t<unk>\t0.0
-99\t<s>\t-0.5
-1.2\t</s>\t0.0
-0.7\tthe\t-0.3
-0.9\tcat\t-0.2
-1.5\tthe cat is not a word\t0.0

\\2-grams:
-0.2\t<s> the\t0.0
-0.3\tthe cat\t0.0
-0.4\tcat </s>\t0.0

\\end\\
"""


def test_arpa_back_off_and_rescoring(tmp_path):
    from speechless_b200.language_model import ArpaLanguageModel, NBestRescorer, find_arpa_file, LN10
    (tmp_path / "tiny.arpa").write_text(ARPA.replace("-1.5\tthe cat is not a word\t0.0\n", "-1.5\tdog\t0.0\n"), encoding="utf8")
    lm = ArpaLanguageModel.read(find_arpa_file(tmp_path))
    assert lm.order == 2
    # seen bigrams: direct look-up; unseen: back-off weight of the history + unigram
    assert lm.log10_probability("cat", ["<s>", "the"]) == pytest.approx(-0.3)
    assert lm.log10_probability("the", ["cat"]) == pytest.approx(-0.2 + -0.7)
    assert lm.log10_probability("zebra", ["the"]) == pytest.approx(-0.3 + -1.0)  # -> <unk>
    assert lm.log10_sentence(["the", "cat"]) == pytest.approx(-0.2 - 0.3 - 0.4)
    rescorer = NBestRescorer(lm)  # reference weights .8 / 0 / 2.3 (net.py:448-451)
    expected = -5.0 + .8 * LN10 * (-0.9) + 2.3 * 2
    assert rescorer.score("the cat", -5.0) == pytest.approx(expected)
    # the language model overrules a slightly better acoustic score of a non-word
    assert rescorer.best([("the cxt", -4.0), ("the cat", -5.0)])[0] == "the cat"


def test_arpa_reader_accepts_space_separated_files_and_rejects_garbage(tmp_path):
    from speechless_b200.language_model import ArpaLanguageModel, find_arpa_file
    spaced = ARPA.replace("-1.5\tthe cat is not a word\t0.0\n", "").replace("\t", " ")
    (tmp_path / "spaced.arpa").write_text(spaced, encoding="utf8")
    lm = ArpaLanguageModel.read(tmp_path / "spaced.arpa")
    assert lm.knows("cat") and not lm.knows("dog")
    assert lm.log10_probability("cat", ["the"]) == pytest.approx(-0.3)
    (tmp_path / "bad.arpa").write_text("\\data\\\n\n\\1-grams:\nnot-a-number\n\\end\\\n", encoding="utf8")
    with pytest.raises(ValueError):
        ArpaLanguageModel.read(tmp_path / "bad.arpa")
    (tmp_path / "empty.arpa").write_text("\\data\\\n\\end\\\n", encoding="utf8")
    with pytest.raises(ValueError):
        ArpaLanguageModel.read(tmp_path / "empty.arpa")
    assert find_arpa_file(tmp_path).name == "bad.arpa"  # (first in sorted order)
    assert find_arpa_file(tmp_path / "nowhere") is None


def _random_case(rng, B, T, V, peaky):
    logits = rng.normal(size=(B, T, V)) * peaky
    probabilities = np.exp(bso.log_softmax(logits)).astype(np.float32)
    lengths = rng.integers(max(1, T // 2), T + 1, size=B).astype(np.int32)
    lengths[0] = T
    return probabilities, lengths


@pytest.mark.gpu
@pytest.mark.parametrize("tf_exact", [True, False])
@pytest.mark.parametrize("B,T,V,beam_width,top_paths,merge,peaky", [
    (1, 5, 2, 1, 1, True, 1.0),
    (3, 40, 5, 1, 1, False, 2.0),
    (4, 60, 6, 4, 3, False, 2.0),      # narrow beam: prefixes drop out and come back
    (4, 60, 6, 4, 3, True, 3.0),
    (6, 80, 5, 2, 2, False, 1.5),      # beam of two, flat distributions: TF's "deactivate child" rule matters
    (6, 80, 5, 3, 3, False, 1.0),
    (3, 120, 29, 16, 4, False, 3.0),
    (2, 200, 29, 100, 8, False, 4.0),  # TF's default width, the English alphabet
    (2, 90, 33, 100, 2, False, 4.0),   # German alphabet
])
def test_cuda_beam_search_matches_the_oracle(monkeypatch, B, T, V, beam_width, top_paths, merge, peaky, tf_exact):
    """tf_exact (the default): TensorFlow's sequential child loop with its order-dependent "deactivate child"
    side effect, against the oracle's literal restatement of ctc_beam_search.h; otherwise
    (SL_BEAM_ORDER_INDEPENDENT=1) the parallel "W best of the union" rule against the oracle's same rule."""
    import torch
    from speechless_b200 import _lib
    lib = _lib.load()
    monkeypatch.setenv("SL_BEAM_ORDER_INDEPENDENT", "0" if tf_exact else "1")
    rng = np.random.default_rng(B * 1000 + T + V + beam_width)
    probabilities, lengths = _random_case(rng, B, T, V, peaky)
    if T == 5:  # the reference's own vector
        probabilities = np.exp(bso.log_softmax(AA_BLANK_AA.astype(np.float64)))[None].astype(np.float32)
        lengths = np.array([5], dtype=np.int32)
    device = torch.device("cuda:0")
    p = torch.from_numpy(probabilities).to(device)
    n = torch.from_numpy(lengths).to(device)
    out = torch.empty((B, top_paths, T), dtype=torch.int32, device=device)
    out_len = torch.empty((B, top_paths), dtype=torch.int32, device=device)
    out_logp = torch.empty((B, top_paths), dtype=torch.float32, device=device)
    ws = torch.empty(lib.sl_ctc_beam_search_workspace_bytes(B, T, beam_width), dtype=torch.uint8, device=device)
    _lib.check(lib.sl_ctc_beam_search_decode(_lib.ptr(p), _lib.ptr(n), _lib.ptr(out), _lib.ptr(out_len),
                                             _lib.ptr(out_logp), B, T, V, V - 1, beam_width, top_paths,
                                             1 if merge else 0, 1, _lib.ptr(ws), ws.numel(), None))
    torch.cuda.synchronize()
    out, out_len, out_logp = out.cpu().numpy(), out_len.cpu().numpy(), out_logp.cpu().numpy()
    for b in range(B):
        scores = np.log(probabilities[b, :lengths[b]].astype(np.float64) + 1e-8)  # net.py:430
        want = bso.beam_search_decode(scores, beam_width=beam_width, top_paths=top_paths, merge_repeated=merge,
                                      tf_deactivation=tf_exact)
        for path, (labels, log_probability) in enumerate(want):
            got = out[b, path, :out_len[b, path]].tolist()
            # fp32 on the device vs fp64 in the oracle: hypotheses closer than 1e-3 nats may swap ranks
            if got != labels:
                alternatives = [lp for lab, lp in want if lab == got]
                assert alternatives and abs(alternatives[0] - log_probability) < 1e-3, (b, path, got, labels)
            else:
                assert out_logp[b, path] == pytest.approx(log_probability, rel=1e-4, abs=1e-3)
            assert (out[b, path, out_len[b, path]:] == -1).all()


# ---------------------------------------------------------------------------------------------------------
# word n-gram model inside the search
# ---------------------------------------------------------------------------------------------------------
def _toy_language_model(rng, letters, order=3, n_words=12, with_unk=True):
    """Random ARPA-style model over words typed with `letters` (plus, sometimes, a word that cannot be typed)."""
    words = set()
    n_words = min(n_words, sum(len(letters) ** k for k in (1, 2, 3)) // 2)  # (there are only so many short words)
    while len(words) < n_words:
        words.add("".join(rng.choice(letters, size=rng.integers(1, 4))))
    words = sorted(words) + ["zq"]  # not typeable: in the model, never in the trie
    ngrams = {("<s>",): (-99.0, float(-rng.random())), ("</s>",): (float(-rng.random() * 2 - 0.2), 0.0)}
    if with_unk:
        ngrams[("<unk>",)] = (float(-rng.random() - 2.0), float(-rng.random() * 0.3))
    for w in words:
        ngrams[(w,)] = (float(-rng.random() * 2 - 0.3), float(-rng.random() * 0.5))
    pool = words + ["</s>"] + (["<unk>"] if with_unk else [])
    for n in range(2, order + 1):
        for _ in range(3 * n_words):
            key = tuple(rng.choice(["<s>"] + words + (["<unk>"] if with_unk else []), size=n - 1)) + (str(rng.choice(pool)),)
            key = tuple(str(k) for k in key)
            if "</s>" in key[:-1] or "<s>" in key[1:]:
                continue
            ngrams[key] = (float(-rng.random() * 1.5 - 0.05), float(-rng.random() * 0.4) if n < order else 0.0)
    return ngrams


def _table_lookup(tables, ids):
    """csrc/beam.cu: lm_find, in numpy."""
    from speechless_b200.language_model import fnv1a
    size = tables.ngrams.shape[0]
    slot = int(fnv1a(np.asarray([ids], dtype=np.int32))[0]) & (size - 1)
    while True:
        row = tables.ngrams[slot]
        if row[0] == 0:
            return None
        if row[0] == len(ids) and row[1:1 + len(ids)].tolist() == list(ids):
            return tuple(row[6:8].view(np.float32).tolist())
        slot = (slot + 1) & (size - 1)


def test_language_model_tables_hold_every_ngram_and_the_vocabulary_trie():
    from speechless_b200.language_model import ArpaLanguageModel, LanguageModelTables
    rng = np.random.default_rng(3)
    alphabet = ["a", "b", "c", " ", "'"]
    ngrams = _toy_language_model(rng, ["a", "b", "c", "'"], order=3, n_words=40)
    model = ArpaLanguageModel(ngrams, 3)
    tables = LanguageModelTables(model, alphabet, symbol_count=len(alphabet) + 1)
    for key, (log10_p, backoff) in ngrams.items():
        got = _table_lookup(tables, [tables.word_id[w] for w in key])
        assert got == pytest.approx((log10_p, backoff), rel=1e-6), key
    assert _table_lookup(tables, [tables.word_id["<s>"], tables.word_id["</s>"], 12345]) is None
    label_of = {c: i for i, c in enumerate(alphabet)}
    for (w,), (log10_p, _) in ((k, v) for k, v in ngrams.items() if len(k) == 1):
        if w in ("<s>", "</s>", "<unk>", "zq"):
            continue
        node = 0
        for c in w:
            node = tables.trie_children[node, label_of[c]]
            assert node > 0 and tables.trie_min_unigram[node] <= np.float32(log10_p)
        assert tables.trie_word[node] == tables.word_id[w]
    assert (tables.trie_children[:, label_of[" "]] == -1).all() and (tables.trie_children[:, -1] == -1).all()
    assert tables.trie_word.max() < len(tables.word_id) and (tables.trie_word >= 0).sum() == 40  # (4 letters: 84 possible)


def test_language_model_tables_are_cached_next_to_the_arpa_file(tmp_path, monkeypatch):
    from speechless_b200 import language_model
    from speechless_b200.language_model import ArpaLanguageModel, LanguageModelTables
    arpa = tmp_path / "tiny.arpa"
    arpa.write_text(ARPA.replace("-1.5\tthe cat is not a word\t0.0\n", "-1.5\tdog\t0.0\n"), encoding="utf8")
    alphabet = list("abcdefghijklmnopqrstuvwxyz '")
    built = LanguageModelTables.from_arpa_file(arpa, alphabet, len(alphabet) + 1)
    assert (tmp_path / "tiny.arpa.sl_tables.npz").exists()

    def no_parsing(path):
        raise AssertionError("the ARPA file was parsed again")
    monkeypatch.setattr(ArpaLanguageModel, "read", staticmethod(no_parsing))
    cached = LanguageModelTables.from_arpa_file(arpa, alphabet, len(alphabet) + 1)
    for name in ("trie_children", "trie_word", "trie_min_unigram", "ngrams"):
        np.testing.assert_array_equal(getattr(cached, name), getattr(built, name))
    for name in LanguageModelTables._SCALARS:
        assert getattr(cached, name) == getattr(built, name), name
    assert language_model.find_arpa_file(tmp_path) == arpa  # the cache is not mistaken for a model
    # another alphabet (or a changed file) invalidates the cache
    with pytest.raises(AssertionError, match="parsed again"):
        LanguageModelTables.from_arpa_file(arpa, alphabet[:-1] + ["-"], len(alphabet) + 1)


def test_oracle_language_model_search_is_exact_with_a_wide_beam():
    """The scorer's deltas telescope: a finished hypothesis holds its CTC probability plus the rescoring formula,
    so a beam wide enough for every prefix returns the arg max of that sum over ALL labelings."""
    from speechless_b200.language_model import ArpaLanguageModel, NBestRescorer
    rng = np.random.default_rng(11)
    alphabet = ["a", "b", " "]
    for case in range(6):
        ngrams = _toy_language_model(rng, ["a", "b"], order=1 + case % 3, n_words=5, with_unk=case % 2 == 0)
        scorer = bso.WordLanguageModelScorer(bso.BackOffModel(ngrams), alphabet)
        rescorer = NBestRescorer(ArpaLanguageModel(ngrams, 1 + case % 3))
        logits = rng.normal(size=(6, 4)) * 2
        totals = bso.brute_force_labelings(logits)
        text = lambda labels: "".join(alphabet[c] for c in labels)
        ranked = sorted(((lp + scorer.sentence_score(text(k)), k) for k, lp in totals.items()), reverse=True)
        got = bso.beam_search_decode(logits, beam_width=10 ** 6, top_paths=3, merge_repeated=False, scorer=scorer)
        for (labels, total), (want_total, want_labels) in zip(got, ranked):
            assert tuple(labels) == want_labels
            assert total == pytest.approx(want_total, abs=1e-9)
            # ... and the host rescorer's formula is the same one (words are split on single spaces there too
            # unless the text has leading / doubled spaces, which str.split() swallows)
            if "  " not in text(labels) and text(labels).strip() == text(labels):
                assert rescorer.score(text(labels), totals[tuple(labels)]) == pytest.approx(total, abs=1e-9)


def _word_like_case(rng, B, T, alphabet, words, peak):
    """Frames that spell sentences made of `words` (one symbol held for 1-3 frames, blanks in between) under
    noise strong enough for the acoustic arg max to be wrong here and there — what a language model is for."""
    V = len(alphabet) + 1
    logits = rng.normal(size=(B, T, V)) * 1.5
    for b in range(B):
        t = 0
        while t < T:
            for c in str(rng.choice(words)) + " ":
                hold = int(rng.integers(1, 4))
                logits[b, t:t + hold, alphabet.index(c)] += peak
                t += hold
                gap = int(rng.integers(0, 3))
                logits[b, t:t + gap, V - 1] += peak
                t += gap
                if t >= T:
                    break
    probabilities = np.exp(bso.log_softmax(logits)).astype(np.float32)
    lengths = rng.integers(max(1, T // 2), T + 1, size=B).astype(np.int32)
    lengths[0] = T
    return probabilities, lengths


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,letters,order,beam_width,top_paths,peaky,with_unk", [
    (3, 30, "ab", 2, 1, 1, 2.0, True),
    (4, 50, "abc", 3, 4, 3, 2.0, True),       # narrow beam, trigram back-off
    (4, 50, "abc", 3, 8, 4, 1.0, False),      # flat distributions, a model without <unk>
    (3, 60, "abcd'", 1, 16, 4, 3.0, True),    # unigram model: only the bonuses and the look-ahead act
    (2, 120, "abcdefghijklmnopqrstuvwxyz'", 3, 100, 8, 4.5, True),  # English alphabet, TF's default width
    (2, 80, "abcdefghijklmnopqrstuvwxyz'", 5, 32, 2, 4.0, True),    # 5-gram model
])
def test_cuda_beam_search_with_language_model_matches_the_oracle(monkeypatch, B, T, letters, order, beam_width,
                                                                 top_paths, peaky, with_unk):
    """sl_ctc_beam_search_decode_lm against the oracle's TF decoder with the word-model scorer hooked in (TF's
    sequential selection; the order-independent rule is not offered with a language model)."""
    tf_exact = True
    import ctypes
    import torch
    from speechless_b200 import _lib
    from speechless_b200.language_model import ArpaLanguageModel, DeviceLanguageModel
    lib = _lib.load()
    monkeypatch.setenv("SL_BEAM_ORDER_INDEPENDENT", "1")  # (ignored with a language model)
    rng = np.random.default_rng(T * 7 + order + beam_width)
    alphabet = list(letters[:len(letters) // 2]) + [" "] + list(letters[len(letters) // 2:])  # space mid-alphabet
    V = len(alphabet) + 1
    ngrams = _toy_language_model(rng, list(letters), order=order, n_words=30, with_unk=with_unk)
    if len(letters) > 20:  # the English-alphabet cases spell sentences of the model's words
        words = [k[0] for k in ngrams if len(k) == 1 and not k[0].startswith("<") and k[0] != "zq"]
        probabilities, lengths = _word_like_case(rng, B, T, alphabet, words, peaky)
    else:
        probabilities, lengths = _random_case(rng, B, T, V, peaky)
    device = torch.device("cuda:0")
    lm = DeviceLanguageModel(ArpaLanguageModel(ngrams, order), alphabet, V, device, kenlm_weight=.8, word_count_weight=.3,
                             valid_word_count_weight=2.3)
    scorer = bso.WordLanguageModelScorer(bso.BackOffModel(ngrams), alphabet, kenlm_weight=.8, word_count_weight=.3,
                                         valid_word_count_weight=2.3)
    p = torch.from_numpy(probabilities).to(device)
    n = torch.from_numpy(lengths).to(device)
    out = torch.empty((B, top_paths, T), dtype=torch.int32, device=device)
    out_len = torch.empty((B, top_paths), dtype=torch.int32, device=device)
    out_logp = torch.empty((B, top_paths), dtype=torch.float32, device=device)
    ws = torch.empty(lib.sl_ctc_beam_search_workspace_bytes(B, T, beam_width), dtype=torch.uint8, device=device)
    _lib.check(lib.sl_ctc_beam_search_decode_lm(_lib.ptr(p), _lib.ptr(n), _lib.ptr(out), _lib.ptr(out_len),
                                                _lib.ptr(out_logp), B, T, V, V - 1, beam_width, top_paths, 0, 1,
                                                ctypes.addressof(lm.struct), _lib.ptr(ws), ws.numel(), None))
    torch.cuda.synchronize()
    out, out_len, out_logp = out.cpu().numpy(), out_len.cpu().numpy(), out_logp.cpu().numpy()
    exact_rank = 0
    for b in range(B):
        scores = np.log(probabilities[b, :lengths[b]].astype(np.float64) + 1e-8)
        want = bso.beam_search_decode(scores, beam_width=beam_width, top_paths=top_paths, merge_repeated=False,
                                      tf_deactivation=tf_exact, scorer=scorer)
        for path, (labels, total) in enumerate(want):
            got = out[b, path, :out_len[b, path]].tolist()
            if got != labels:  # fp32 vs fp64: hypotheses closer than 2e-3 may swap ranks
                alternatives = [lp for lab, lp in want if lab == got]
                assert alternatives and abs(alternatives[0] - total) < 2e-3, (b, path, got, labels)
            else:
                exact_rank += 1
                assert out_logp[b, path] == pytest.approx(total, rel=1e-4, abs=2e-3)
    assert exact_rank >= B  # (at the very least every best path)


@pytest.mark.gpu
def test_wav2letter_decodes_with_beam_search_and_language_model(tmp_path):
    from speechless_b200 import english_frequent_characters
    from speechless_b200.net import Wav2Letter
    from speechless_b200.labeled_example import LabeledSpectrogram

    class Example(LabeledSpectrogram):
        def __init__(self, id, label, spectrogram):
            self.id, self.label, self._s = id, label, spectrogram

        def z_normalized_transposed_spectrogram(self):
            return self._s

    (tmp_path / "vocabulary").write_text("".join(english_frequent_characters).upper(), encoding="utf8")
    with pytest.raises(NotImplementedError):  # a KenLM binary alone cannot be read
        Wav2Letter(128, english_frequent_characters, kenlm_directory=tmp_path, main_filter_count=64,
                   out_filter_count=64, device="cuda:0", seed=3)
    (tmp_path / "tiny.arpa").write_text(ARPA.replace("-1.5\tthe cat is not a word\t0.0\n", "-1.5\tdog\t0.0\n"), encoding="utf8")
    net = Wav2Letter(128, english_frequent_characters, kenlm_directory=tmp_path, main_filter_count=64,
                     out_filter_count=64, device="cuda:0", seed=3, decoder_beam_width=16, decoder_top_paths=8,
                     language_model_mode="rescoring")
    in_search = Wav2Letter(128, english_frequent_characters, kenlm_directory=tmp_path, main_filter_count=64,
                           out_filter_count=64, device="cuda:0", seed=3, decoder_beam_width=16)
    assert in_search.language_model_mode == "in-search" and in_search.device_language_model is not None
    greedy = Wav2Letter(128, english_frequent_characters, main_filter_count=64, out_filter_count=64,
                        device="cuda:0", seed=3)
    rng = np.random.default_rng(1)
    batch = [Example("a", "the cat", rng.standard_normal((120, 128))), Example("b", "dog", rng.standard_normal((90, 128)))]
    with_lm = net.test_and_predict_batch(batch)
    without = greedy.test_and_predict_batch(batch)
    # same weights (same seed) -> same losses; predictions are strings over the alphabet
    assert [r.loss for r in with_lm.results] == pytest.approx([r.loss for r in without.results], rel=1e-5)
    assert all(set(r.predicted) <= set(english_frequent_characters) for r in with_lm.results)
    # the winner is one of the n-best hypotheses of the plain device beam search
    tower = net.tower
    ws = tower.upload(net._input_batch_and_prediction_lengths([e.z_normalized_transposed_spectrogram() for e in batch])[0])
    tower.forward(ws)
    tower.set_prediction_lengths(ws, [60, 45])
    n_best = net.beam_search_batch(ws, top_paths=8)
    for result, hypotheses in zip(with_lm.results, n_best):
        texts = [net.grapheme_encoding.decode_graphemes(g, merge_repeated=False) for g, _ in hypotheses]
        assert result.predicted in texts
    # default mode: the language model inside the search — the oracle's TF decoder with the word-model scorer on
    # the device's own probabilities gives the same strings
    from speechless_b200.language_model import ArpaLanguageModel, find_arpa_file
    arpa = ArpaLanguageModel.read(find_arpa_file(tmp_path))
    scorer = bso.WordLanguageModelScorer(bso.BackOffModel(arpa.ngrams), list(english_frequent_characters))
    searched = in_search.test_and_predict_batch(batch)
    probabilities = ws.probs.cpu().numpy()
    for b, (result, frames) in enumerate(zip(searched.results, [60, 45])):
        scores = np.log(probabilities[b, :frames].astype(np.float64) + 1e-8)
        want = bso.beam_search_decode(scores, beam_width=16, top_paths=2, merge_repeated=False, scorer=scorer)
        texts = ["".join(english_frequent_characters[c] for c in labels) for labels, _ in want]
        assert result.predicted == texts[0] or (result.predicted == texts[1] and abs(want[0][1] - want[1][1]) < 2e-3)
    assert [r.loss for r in searched.results] == pytest.approx([r.loss for r in without.results], rel=1e-5)
    spectrograms_ = [e.z_normalized_transposed_spectrogram() for e in batch]
    ranked_lm = in_search.predict_batch_with_beam_search(spectrograms_, beam_width=16, top_paths=3, use_language_model=True)
    assert [h[0][0] for h in ranked_lm] == [r.predicted for r in searched.results]
    assert all([s for _, s in h] == sorted([s for _, s in h], reverse=True) for h in ranked_lm)
    with pytest.raises(ValueError, match="kenlm_directory"):
        greedy.predict_batch_with_beam_search(spectrograms_, use_language_model=True)
    # the public beam-search call: best hypothesis first, log-probabilities descending; with width 1 on
    # these near-uniform outputs it still returns exactly one hypothesis per utterance
    spectrograms = [e.z_normalized_transposed_spectrogram() for e in batch]
    ranked = greedy.predict_batch_with_beam_search(spectrograms, beam_width=16, top_paths=4)
    assert len(ranked) == 2 and all(len(h) == 4 for h in ranked)
    for hypotheses in ranked:
        scores = [score for _, score in hypotheses]
        assert scores == sorted(scores, reverse=True)
    assert [len(h) for h in greedy.predict_batch_with_beam_search(spectrograms, beam_width=1)] == [1, 1]
