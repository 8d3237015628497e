"""Word n-gram language model for the beam-search decoder (host side: ARPA reader, device tables).

The reference decodes with a language model only through a patched TensorFlow
(`tf.nn.ctc_beam_search_decoder(kenlm_directory_path=..., kenlm_weight=.8, word_count_weight=0,
valid_word_count_weight=2.3)`, reference net.py:444-451; fork named in net.py:420-422 and
README.md:17).  Neither that fork nor the KenLM library is in the reference tree or installable
here; what is kept is the interface (`kenlm_directory` with its `vocabulary` file, net.py:171-177) and
the three weights.  A finished hypothesis scores

    score(hypothesis) = log P_ctc + kenlm_weight * ln P_lm(words </s>)
                        + word_count_weight * #words + valid_word_count_weight * #words known to the LM

and there are two ways to get there:
  * IN-SEARCH (default): `DeviceLanguageModel` turns the ARPA model into a vocabulary trie and an n-gram
    hash table in device memory, and `sl_ctc_beam_search_decode_lm` runs TF's scorer hooks on the device
    (csrc/beam.cu) — words are scored when the space label arrives, unfinished words carry a look-ahead;
  * N-BEST RE-SCORING (`NBestRescorer`): the plain device beam search, its finished hypotheses re-ranked
    on the host with the same formula.

**Parity unpinned** against the fork (stated in DESIGN.md).  The model file is the plain-text ARPA
format every KenLM installation can write (`lmplz`); KenLM's binary format is not read.
"""
import ctypes
import math
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

LN10 = math.log(10.0)


class ArpaLanguageModel:
    """Back-off n-gram model read from an ARPA file (log10 probabilities and back-off weights)."""

    def __init__(self, ngrams: Dict[Tuple[str, ...], Tuple[float, float]], order: int):
        self.ngrams = ngrams
        self.order = order
        self.unknown = ngrams.get(("<unk>",), (-100.0, 0.0))[0]

    @staticmethod
    def read(path: Path) -> "ArpaLanguageModel":
        ngrams: Dict[Tuple[str, ...], Tuple[float, float]] = {}
        order, current = 0, 0
        with open(str(path), encoding="utf8") as lines:
            for raw in lines:
                line = raw.strip()
                if not line or line == "\\data\\" or line.startswith("ngram "):
                    continue
                if line == "\\end\\":
                    break
                if line.startswith("\\") and line.endswith("-grams:"):
                    current = int(line[1:line.index("-")])
                    order = max(order, current)
                    continue
                fields = line.split("\t") if "\t" in line else line.split()
                if current == 0 or len(fields) < 2:
                    raise ValueError("Malformed ARPA line: {!r}".format(raw))
                if "\t" in line:
                    words = tuple(fields[1].split(" "))
                    backoff = float(fields[2]) if len(fields) > 2 else 0.0
                else:
                    words = tuple(fields[1:1 + current])
                    backoff = float(fields[1 + current]) if len(fields) > 1 + current else 0.0
                if len(words) != current:
                    raise ValueError("Malformed ARPA line: {!r}".format(raw))
                ngrams[words] = (float(fields[0]), backoff)
        if order == 0:
            raise ValueError("No n-grams found in {}".format(path))
        return ArpaLanguageModel(ngrams, order)

    def knows(self, word: str) -> bool:
        return (word,) in self.ngrams

    def log10_probability(self, word: str, history: Sequence[str]) -> float:
        """log10 P(word | history) with Katz back-off (the ARPA semantics KenLM implements)."""
        if not self.knows(word):
            word = "<unk>"
            if not self.knows(word):
                return self.unknown
        history = tuple(history[-(self.order - 1):]) if self.order > 1 else ()
        penalty = 0.0
        while True:
            entry = self.ngrams.get(history + (word,))
            if entry is not None:
                return penalty + entry[0]
            if not history:
                return penalty + self.unknown
            context = self.ngrams.get(history)
            if context is not None:
                penalty += context[1]
            history = history[1:]

    def log10_sentence(self, words: Sequence[str]) -> float:
        """log10 P(<s> words </s>) (without the probability of <s> itself, as KenLM's `score`)."""
        history: List[str] = ["<s>"]
        total = 0.0
        for word in list(words) + ["</s>"]:
            total += self.log10_probability(word, history)
            history.append(word if self.knows(word) else "<unk>")
        return total


class NBestRescorer:
    """Picks the best of the beam-search hypotheses under CTC + language-model score (weights as in
    reference net.py:448-451)."""

    def __init__(self, language_model: ArpaLanguageModel, kenlm_weight: float = .8, word_count_weight: float = 0.,
                 valid_word_count_weight: float = 2.3):
        self.language_model = language_model
        self.kenlm_weight = kenlm_weight
        self.word_count_weight = word_count_weight
        self.valid_word_count_weight = valid_word_count_weight

    def score(self, text: str, ctc_log_probability: float) -> float:
        words = text.split()
        valid = sum(1 for w in words if self.language_model.knows(w))
        return (ctc_log_probability + self.kenlm_weight * LN10 * self.language_model.log10_sentence(words) +
                self.word_count_weight * len(words) + self.valid_word_count_weight * valid)

    def best(self, hypotheses: Sequence[Tuple[str, float]]) -> Tuple[str, float]:
        """hypotheses: (text, CTC log-probability), any order; ties keep the earlier (better CTC) one."""
        scored = [(self.score(text, log_probability), -index, text) for index, (text, log_probability) in
                  enumerate(hypotheses)]
        best_score, _, best_text = max(scored)
        return best_text, best_score


def find_arpa_file(kenlm_directory: Path) -> Optional[Path]:
    candidates = sorted(Path(kenlm_directory).glob("*.arpa"))
    return candidates[0] if candidates else None


class SlWordLm(ctypes.Structure):
    """include/speechless_b200.h: struct SlWordLm."""
    _fields_ = [("trie_children", ctypes.c_void_p), ("trie_word", ctypes.c_void_p),
                ("trie_min_unigram", ctypes.c_void_p), ("ngrams", ctypes.c_void_p),
                ("ngram_mask", ctypes.c_uint32), ("n_labels", ctypes.c_int32), ("space_label", ctypes.c_int32),
                ("order", ctypes.c_int32), ("bos_id", ctypes.c_int32), ("eos_id", ctypes.c_int32),
                ("unk_id", ctypes.c_int32), ("has_unk", ctypes.c_int32), ("unknown_log10", ctypes.c_float),
                ("weight", ctypes.c_float), ("word_count_weight", ctypes.c_float),
                ("valid_word_count_weight", ctypes.c_float)]


def fnv1a(ids):
    """32-bit FNV-1a over the rows of an (n, k) int array (what csrc/beam.cu: lm_find hashes)."""
    import numpy
    h = numpy.full(len(ids), 2166136261, dtype=numpy.uint64)
    for k in range(ids.shape[1]):
        h = ((h ^ (ids[:, k].astype(numpy.int64) & 0xffffffff).astype(numpy.uint64)) * numpy.uint64(16777619)) \
            & numpy.uint64(0xffffffff)
    return h.astype(numpy.uint32)


class LanguageModelTables:
    """The ARPA model as flat arrays (numpy, host): vocabulary trie over the label alphabet and the n-gram
    open-addressing table.  `alphabet[label]` is the character of a label (the CTC blank has none)."""
    MAX_ORDER = 5

    def __init__(self, model: ArpaLanguageModel, alphabet: Sequence[str], symbol_count: int):
        import numpy
        if model.order > self.MAX_ORDER:
            raise ValueError("n-gram order {} > {}".format(model.order, self.MAX_ORDER))
        if " " not in alphabet:
            raise ValueError("the alphabet has no space: words cannot end")
        self.order, self.symbol_count = model.order, symbol_count
        self.space_label = list(alphabet).index(" ")
        words = sorted({w for key in model.ngrams for w in key})
        self.word_id = {w: i for i, w in enumerate(words)}
        self.has_unk = "<unk>" in self.word_id
        self.unk_id = self.word_id.get("<unk>", len(words))
        self.bos_id = self.word_id.get("<s>", len(words) + 1)
        self.eos_id = self.word_id.get("</s>", self.unk_id)  # (a model without </s> scores the end as an unknown word)
        self.unknown_log10 = model.unknown
        # ---- trie of the words that can be typed with the alphabet
        label_of = {c: i for i, c in enumerate(alphabet) if c != " "}
        children = [[-1] * symbol_count]
        word_at, min_unigram = [-1], [0.0]
        for (w,), (log10_p, _) in ((k, v) for k, v in model.ngrams.items() if len(k) == 1):
            if w in ("<s>", "</s>", "<unk>") or not w or any(c not in label_of for c in w):
                continue
            node = 0
            min_unigram[0] = min(min_unigram[0], log10_p)
            for c in w:
                nxt = children[node][label_of[c]]
                if nxt < 0:
                    nxt = len(children)
                    children[node][label_of[c]] = nxt
                    children.append([-1] * symbol_count)
                    word_at.append(-1)
                    min_unigram.append(0.0)
                node = nxt
                min_unigram[node] = min(min_unigram[node], log10_p)
            word_at[node] = self.word_id[w]
        self.trie_children = numpy.asarray(children, dtype=numpy.int32)
        self.trie_word = numpy.asarray(word_at, dtype=numpy.int32)
        self.trie_min_unigram = numpy.asarray(min_unigram, dtype=numpy.float32)
        # ---- n-gram table: {n, id0..id4, log10 p, log10 back-off}, slot = FNV-1a(ids) & mask, linear probing
        count = len(model.ngrams)
        size = 64
        while size < 2 * count:
            size *= 2
        table = numpy.zeros((size, 8), dtype=numpy.int32)
        keys = numpy.full((count, 5), -1, dtype=numpy.int32)
        lengths = numpy.zeros(count, dtype=numpy.int32)
        values = numpy.zeros((count, 2), dtype=numpy.float32)
        for row, (key, (log10_p, backoff)) in enumerate(model.ngrams.items()):
            lengths[row] = len(key)
            keys[row, :len(key)] = [self.word_id[w] for w in key]
            values[row] = (log10_p, backoff)
        slots = numpy.zeros(count, dtype=numpy.int64)
        for n in range(1, self.order + 1):
            rows = numpy.nonzero(lengths == n)[0]
            if len(rows):
                slots[rows] = fnv1a(keys[rows, :n]) & (size - 1)
        pending = numpy.arange(count)
        while len(pending):  # vectorised linear probing: per round, the first claimant of every free slot wins
            wanted = slots[pending]
            free = table[wanted, 0] == 0
            _, first = numpy.unique(wanted, return_index=True)
            winner = numpy.zeros(len(pending), dtype=bool)
            winner[first] = True
            winner &= free
            rows = pending[winner]
            table[slots[rows], 0] = lengths[rows]
            table[slots[rows], 1:6] = keys[rows]
            table[slots[rows], 6:8] = values[rows].view(numpy.int32)
            pending = pending[~winner]
            slots[pending] = (slots[pending] + 1) & (size - 1)
        self.ngrams = table


    # ---- cache: parsing a large ARPA file and hashing its n-grams in Python takes minutes, loading arrays seconds
    _SCALARS = ("order", "symbol_count", "space_label", "has_unk", "unk_id", "bos_id", "eos_id", "unknown_log10")

    def save(self, path, fingerprint: str) -> None:
        import numpy
        numpy.savez(str(path), trie_children=self.trie_children, trie_word=self.trie_word,
                    trie_min_unigram=self.trie_min_unigram, ngrams=self.ngrams, fingerprint=numpy.array(fingerprint),
                    scalars=numpy.array([float(getattr(self, name)) for name in self._SCALARS], dtype=numpy.float64))

    @classmethod
    def load(cls, path, fingerprint: str) -> Optional["LanguageModelTables"]:
        """The tables saved under `path` if they were built from the same model file and alphabet, else None."""
        import numpy
        try:
            with numpy.load(str(path)) as z:
                if str(z["fingerprint"]) != fingerprint:
                    return None
                tables = cls.__new__(cls)
                tables.trie_children, tables.trie_word = z["trie_children"], z["trie_word"]
                tables.trie_min_unigram, tables.ngrams = z["trie_min_unigram"], z["ngrams"]
                for name, value in zip(cls._SCALARS, z["scalars"]):
                    setattr(tables, name, float(value) if name == "unknown_log10" else int(value))
                tables.has_unk = bool(tables.has_unk)
                tables.word_id = None  # (only needed while building)
                return tables
        except (OSError, KeyError, ValueError):
            return None

    @classmethod
    def from_arpa_file(cls, arpa_file: Path, alphabet: Sequence[str], symbol_count: int,
                       cache: bool = True) -> "LanguageModelTables":
        """Tables of the model in `arpa_file`; cached next to it as `<name>.sl_tables.npz`, keyed by the file's size
        and modification time and by the alphabet (a stale or unwritable cache is simply rebuilt / skipped)."""
        arpa_file = Path(arpa_file)
        stat = arpa_file.stat()
        fingerprint = "{}:{}:{}:{}:{}".format(arpa_file.name, stat.st_size, int(stat.st_mtime), symbol_count,
                                              "".join(alphabet))
        cache_path = arpa_file.with_name(arpa_file.name + ".sl_tables.npz")
        if cache and cache_path.exists():
            tables = cls.load(cache_path, fingerprint)
            if tables is not None:
                return tables
        tables = cls(ArpaLanguageModel.read(arpa_file), alphabet, symbol_count)
        if cache:
            try:
                tables.save(cache_path, fingerprint)
            except OSError:
                pass
        return tables


class DeviceLanguageModel:
    """`LanguageModelTables` in device memory plus the `SlWordLm` struct that points at them.  `model`: an
    `ArpaLanguageModel`, or ready-made `LanguageModelTables` (e.g. `LanguageModelTables.from_arpa_file`)."""

    def __init__(self, model, alphabet: Sequence[str], symbol_count: int, device,
                 kenlm_weight: float = .8, word_count_weight: float = 0., valid_word_count_weight: float = 2.3):
        import torch
        tables = model if isinstance(model, LanguageModelTables) else LanguageModelTables(model, alphabet, symbol_count)
        if tables.symbol_count != symbol_count or tables.space_label != list(alphabet).index(" "):
            raise ValueError("the language-model tables were built for another alphabet")
        self.tables = tables
        self._tensors = [torch.from_numpy(a).to(device).contiguous() for a in
                         (tables.trie_children, tables.trie_word, tables.trie_min_unigram, tables.ngrams)]
        self.struct = SlWordLm(
            trie_children=self._tensors[0].data_ptr(), trie_word=self._tensors[1].data_ptr(),
            trie_min_unigram=self._tensors[2].data_ptr(), ngrams=self._tensors[3].data_ptr(),
            ngram_mask=tables.ngrams.shape[0] - 1, n_labels=symbol_count, space_label=tables.space_label,
            order=tables.order, bos_id=tables.bos_id, eos_id=tables.eos_id, unk_id=tables.unk_id,
            has_unk=1 if tables.has_unk else 0, unknown_log10=tables.unknown_log10, weight=kenlm_weight * LN10,
            word_count_weight=word_count_weight, valid_word_count_weight=valid_word_count_weight)
