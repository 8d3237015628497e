"""`Wav2Letter` — the reference's model/loss/decode/train surface on hand-written sm_100a kernels.

Host-side mirror of `speechless/net.py::Wav2Letter` (reference net.py:117-607): same class,
method and argument names, return types, log lines and error behaviour, so
`Configuration.train` (configuration.py:96-101), `test_model` (:124-125), `load_model`
(:168-178), `main.py:144,202-206` and the README snippets call it unchanged.  Below the
surface nothing is Keras/TensorFlow: `ConvTower` (engine.py) drives libspeechless_b200.so.

Differences that are deliberate and documented in DESIGN.md:
* `optimizer` defaults to a fresh `Adam(1e-4)` per instance (the reference shares one
  default-argument instance between all models, net.py:132);
* batches of size 1 work (the reference needs the duplication workaround of net.py:492-495,
  which is kept so results are identical either way);
* weights are saved as `.npz` keyed by the Keras layer names and, for `.h5` paths, in Keras' HDF5 layout
  (h5py when installed, else the pure-Python `hdf5_lite`); `.h5` files written by Keras load the same way;
* `use_raw_wave_input=True` adds `wave_conv` (k250 s160, net.py:310-312) in front; its receptive
  fields are laid out as rows on the device so it runs on the same tensor-core kernels;
* `kenlm_directory` keeps the vocabulary check of net.py:171-177 and decodes with the device prefix
  beam search with a word n-gram model inside the search (`sl_ctc_beam_search_decode_lm`; the model is read
  from a `*.arpa` file of that directory, language_model.py; `language_model_mode="rescoring"` re-ranks the
  n-best list of the plain search instead).  The scorer of the reference's patched TensorFlow (net.py:420-422)
  is not in the tree, so parity with it is unpinned; a KenLM binary alone raises `NotImplementedError`;
  `use_asg=True` raises at loss time exactly like the reference;
* `dropout` masks come from a counter-based hash, so they cannot equal TF's bit for bit; the
  arithmetic given the masks is what the parity tests check.
"""
import itertools
import json
import queue
import threading
import time
from collections import OrderedDict
from functools import reduce
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import numpy
from numpy import ndarray, zeros, array, reshape, concatenate

from speechless_b200._lib import PRECISIONS
from speechless_b200.grapheme_enconding import CtcGraphemeEncoding, AsgGraphemeEncoding
from speechless_b200.labeled_example import LabeledSpectrogram
from speechless_b200.results import (ExpectationVsPrediction, ExpectationsVsPredictions,
                                     ExpectationsVsPredictionsInBatches, ExpectationsVsPredictionsInGroupedBatches)
from speechless_b200.tools import log, mkdir, read_text, single, single_or_none

__all__ = ["Adam", "Wav2Letter", "ExpectationVsPrediction", "ExpectationsVsPredictions",
           "ExpectationsVsPredictionsInBatches", "ExpectationsVsPredictionsInGroupedBatches"]


class Adam:
    """Stands in for `keras.optimizers.Adam` (reference net.py:12,132): Keras-2 update rule,
    hyper-parameter names as in Keras.  The state (moments) lives on the device in `ConvTower`."""

    def __init__(self, lr: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-8):
        self.lr = lr
        self.beta_1 = beta_1
        self.beta_2 = beta_2
        self.epsilon = epsilon
        self.iterations = 0


class _ConvLayerView:
    """What callers of `predictive_net.layers[i]` use: name, strides, trainable, get/set_weights."""

    def __init__(self, tower, index: int):
        self._tower = tower
        self._index = index
        spec = tower.layers[index]
        self.name = spec.name
        self.filters = spec.cout
        self.kernel_size = (spec.kernel,)
        self.strides = (spec.stride,)
        self.trainable = index >= tower.frozen_layer_count

    def get_weights(self) -> List[ndarray]:
        """[kernel (k, C_in, C_out), bias (C_out,)] — the Keras layout (reference net.py:238,251-255)."""
        return self._tower.get_layer_weights(self._index)

    def set_weights(self, weights: List[ndarray]) -> None:
        kernel, bias = weights
        self._tower.set_layer_weights(self._index, kernel, bias)


class _DropoutLayerView:
    """Stands for the `keras.layers.Dropout` the reference inserts in front of striding_conv and
    inner_conv_1..7 (net.py:301-303); it counts as a layer for `frozen_layer_count` (net.py:338-339)."""

    def __init__(self, name: str, rate: float):
        self.name = name
        self.rate = rate
        self.trainable = True

    def get_weights(self) -> List[ndarray]:
        return []

    def set_weights(self, weights: List[ndarray]) -> None:
        if weights:
            raise ValueError("Dropout layers have no weights")


class PredictiveNet:
    """Shim for the Keras `Sequential` the reference exposes as `Wav2Letter.predictive_net`
    (net.py:168, used by main.py:121 and load_weights :212,237-269)."""

    def __init__(self, tower, input_size_per_time_step: int):
        self._tower = tower
        self.conv_layers = [_ConvLayerView(tower, index) for index in range(len(tower.layers))]
        # the reference's layer list interleaves Dropout layers when dropout is configured
        self.layers = []
        for index, view in enumerate(self.conv_layers):
            if index in tower.dropout_layers:
                self.layers.append(_DropoutLayerView("dropout_before_{}".format(view.name), tower.dropout))
            self.layers.append(view)
        self.input_shape = (None, None, input_size_per_time_step)

    @staticmethod
    def _npz_path(path) -> Path:
        path = Path(str(path))
        return path if path.suffix == ".npz" else path.with_suffix(".npz")

    def save_weights(self, path) -> None:
        """`.npz` keyed by the Keras layer names, or — a path ending in `.h5` / `.hdf5` — the HDF5 layout of
        Keras' `save_weights` (root attribute `layer_names`, one group per layer with `weight_names` and the
        datasets `<layer>/kernel:0`, `<layer>/bias:0`, reference net.py:572), written by h5py when it is
        installed and by `speechless_b200.hdf5_lite` otherwise; next to an `.h5` the `.npz` is always written too
        (hdf5_lite's writer could not be checked against libhdf5 in this image)."""
        arrays = {}
        for layer in self.conv_layers:
            kernel, bias = layer.get_weights()
            arrays["{}/kernel".format(layer.name)] = kernel
            arrays["{}/bias".format(layer.name)] = bias
        numpy.savez(str(self._npz_path(path)), **arrays)
        if Path(str(path)).suffix not in (".h5", ".hdf5"):
            return
        layer_names = [layer.name.encode("utf8") for layer in self.conv_layers]
        weight_names = {layer.name: ["{}/kernel:0".format(layer.name), "{}/bias:0".format(layer.name)]
                        for layer in self.conv_layers}
        try:
            import h5py  # optional
        except ImportError:
            from speechless_b200 import hdf5_lite
            tree = {layer.name: {weight_names[layer.name][0]: arrays["{}/kernel".format(layer.name)],
                                 weight_names[layer.name][1]: arrays["{}/bias".format(layer.name)]}
                    for layer in self.conv_layers}
            attrs = {"/": {"layer_names": layer_names, "backend": b"tensorflow", "keras_version": b"2.0.4"}}
            for name, names in weight_names.items():
                attrs[name] = {"weight_names": [n.encode("utf8") for n in names]}
            hdf5_lite.write(path, tree, attrs)
            return
        with h5py.File(str(path), "w") as f:
            f.attrs["layer_names"] = layer_names
            for layer in self.conv_layers:
                group = f.create_group(layer.name)
                names = weight_names[layer.name]
                group.attrs["weight_names"] = [n.encode("utf8") for n in names]
                group.create_dataset(names[0], data=arrays["{}/kernel".format(layer.name)])
                group.create_dataset(names[1], data=arrays["{}/bias".format(layer.name)])

    def load_weights(self, path) -> None:
        """The exact file given wins: an `.h5` written by Keras (h5py's default file format) is read with h5py
        when it is installed, else with `speechless_b200.hdf5_lite` (pure Python; validated on a genuine
        libhdf5-written file, tests/test_hdf5_lite.py); `weights-epochN.npz` next to it is the fallback when
        the `.h5` does not exist."""
        npz = self._npz_path(path)
        exact = Path(str(path))
        if npz.exists() and (exact == npz or not exact.exists()):
            with numpy.load(str(npz)) as arrays:
                for layer in self.conv_layers:
                    layer.set_weights([arrays["{}/kernel".format(layer.name)], arrays["{}/bias".format(layer.name)]])
            return
        try:
            import h5py
            opened = h5py.File(str(path), "r")
        except ImportError:
            from speechless_b200 import hdf5_lite
            opened = hdf5_lite.File(path)
        with opened as f:
            root = f["model_weights"] if "model_weights" in f else f
            for layer in self.conv_layers:
                group = root[layer.name]
                names = [n.decode("utf8") if isinstance(n, bytes) else n for n in group.attrs["weight_names"]]
                # Keras stores (kernel, bias) in `weight_names` order; a Conv1D kernel is (k, C_in, C_out)
                layer.set_weights([numpy.asarray(group[names[0]]), numpy.asarray(group[names[1]])])


class _Prefetcher:
    """Keras' fit_generator pulls batches on a background thread with a bounded queue
    (net.py:550); this keeps host batch assembly overlapped with the device step."""
    _END = object()

    def __init__(self, iterable: Iterable, depth: int = 10):
        self._queue: "queue.Queue" = queue.Queue(maxsize=depth)
        self._error: Optional[BaseException] = None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, args=(iter(iterable),), daemon=True)
        self._thread.start()

    def _run(self, iterator):
        try:
            for item in iterator:
                while not self._stop.is_set():
                    try:
                        self._queue.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if self._stop.is_set():
                    return
        except BaseException as e:  # surfaced on the consumer side
            self._error = e
        self._queue.put(self._END)

    def __iter__(self):
        return self

    def __next__(self):
        item = self._queue.get()
        if item is self._END:
            self._queue.put(self._END)
            if self._error is not None:
                raise self._error
            raise StopIteration
        return item

    def close(self):
        self._stop.set()


class Wav2Letter:
    """Speech-recognition network based on wav2letter (https://arxiv.org/pdf/1609.03193v2.pdf)."""

    class InputNames:
        input_batch = "input_batch"
        label_batch = "label_batch"
        prediction_lengths = "prediction_lenghts"
        label_lengths = "label_lenghts"

    def __init__(self,
                 input_size_per_time_step: int,
                 allowed_characters: List[chr],
                 use_raw_wave_input: bool = False,
                 activation: str = "relu",
                 output_activation: str = "softmax",
                 optimizer: Optional[Adam] = None,
                 dropout: Optional[float] = None,
                 load_model_from_directory: Optional[Path] = None,
                 load_epoch: Optional[int] = None,
                 allowed_characters_for_loaded_model: Optional[List[chr]] = None,
                 frozen_layer_count: int = 0,
                 reinitialize_trainable_loaded_layers: bool = False,
                 use_asg: bool = False,
                 asg_transition_probabilities: Optional[ndarray] = None,
                 asg_initial_probabilities: Optional[ndarray] = None,
                 kenlm_directory: Path = None,
                 *,
                 main_filter_count: int = 250,
                 out_filter_count: int = 2000,
                 compute_dtype: str = "bf16x2",
                 device=None,
                 seed: Optional[int] = None,
                 decoder_beam_width: int = 100,
                 decoder_top_paths: int = 32,
                 language_model_mode: str = "in-search",
                 data_parallel=None,
                 devices=None):
        if frozen_layer_count > 0 and load_model_from_directory is None:
            raise ValueError("Layers cannot be frozen if model is trained from scratch.")
        if compute_dtype not in PRECISIONS:
            raise ValueError("compute_dtype must be one of {}".format(sorted(PRECISIONS)))

        self.kenlm_directory = kenlm_directory
        self.rescorer = None
        self.device_language_model = None
        if language_model_mode not in ("in-search", "rescoring"):
            raise ValueError("language_model_mode must be 'in-search' or 'rescoring'")
        self.language_model_mode = language_model_mode
        self.decoder_beam_width = decoder_beam_width
        self.decoder_top_paths = min(decoder_top_paths, decoder_beam_width)
        self.grapheme_encoding = AsgGraphemeEncoding(allowed_characters=allowed_characters) \
            if use_asg else CtcGraphemeEncoding(allowed_characters=allowed_characters)
        self.asg_transition_probabilities = asg_transition_probabilities
        self.asg_initial_probabilities = asg_initial_probabilities
        self.use_asg = use_asg
        self.frozen_layer_count = frozen_layer_count
        self.output_activation = output_activation
        self.activation = activation
        self.use_raw_wave_input = use_raw_wave_input
        self.input_size_per_time_step = input_size_per_time_step
        self.optimizer = optimizer if optimizer is not None else Adam(1e-4)
        self.load_epoch = load_epoch
        self.dropout = dropout
        self.main_filter_count = main_filter_count
        self.out_filter_count = out_filter_count
        self.compute_dtype = compute_dtype
        self.seed = seed
        # One process per GPU: `data_parallel` (speechless_b200.distributed.DataParallel) makes
        # `train` shard every batch by utterance and `test_and_predict_batches` shard by batch; the device
        # defaults to the rank's own GPU and rank 0's initial weights are broadcast to every replica.
        self.data_parallel = data_parallel if (data_parallel is not None and data_parallel.active) else None
        # `devices`: the GPUs of the job.  One entry = `device`; several entries name the GPUs of a
        # data-parallel job, one process per GPU — this process takes the entry of its local rank.
        if devices is not None:
            devices = list(devices)
            if device is not None or not devices:
                raise ValueError("pass either device or a non-empty devices list")
            if len(devices) == 1:
                device = devices[0]
            elif self.data_parallel is None or self.data_parallel.world_size != len(devices):
                raise ValueError("{} devices need one process per GPU: launch with torchrun --nproc-per-node {} and "
                                 "pass data_parallel=speechless_b200.distributed.DataParallel()".format(
                                     len(devices), len(devices)))
            else:
                device = devices[self.data_parallel.local_rank]
            if isinstance(device, int):
                device = "cuda:{}".format(device)
        if device is None and self.data_parallel is not None:
            device = "cuda:{}".format(self.data_parallel.local_rank)
        self._device = device
        self.predictive_net = self.create_predictive_net()
        self.prediction_phase_flag = 0.

        if self.kenlm_directory is not None:
            expected_characters = list(
                single(read_text(self.kenlm_directory / "vocabulary", encoding='utf8').splitlines()).lower())
            if allowed_characters != expected_characters:
                raise ValueError("Allowed characters {} differ from those expected by kenlm decoder: {}".
                                 format(allowed_characters, expected_characters))
            # The reference hands the directory to a patched TensorFlow whose beam search scores words with
            # KenLM (net.py:420-422,444-451).  Here: a word n-gram model in ARPA text format with the same three
            # weights, scored INSIDE the device beam search (`language_model_mode="in-search"`, default) or used
            # to re-rank the finished hypotheses of the plain search ("rescoring") — language_model.py.
            from speechless_b200.language_model import (ArpaLanguageModel, DeviceLanguageModel, NBestRescorer,
                                                        find_arpa_file)
            arpa_file = find_arpa_file(self.kenlm_directory)
            if arpa_file is None:
                raise NotImplementedError(
                    "No *.arpa file in {}: KenLM's binary format needs the KenLM library of the reference's patched "
                    "TensorFlow (net.py:420-422), which is not available; export the model as ARPA text.".format(
                        self.kenlm_directory))
            if language_model_mode == "in-search":
                # (device tables are cached next to the ARPA file: a large model is parsed once)
                from speechless_b200.language_model import LanguageModelTables
                symbol_count = self.grapheme_encoding.grapheme_set_size
                self.device_language_model = DeviceLanguageModel(
                    LanguageModelTables.from_arpa_file(arpa_file, allowed_characters, symbol_count),
                    alphabet=allowed_characters, symbol_count=symbol_count, device=self.tower.device,
                    kenlm_weight=.8, word_count_weight=0, valid_word_count_weight=2.3)
            else:
                self.rescorer = NBestRescorer(ArpaLanguageModel.read(arpa_file), kenlm_weight=.8, word_count_weight=0,
                                              valid_word_count_weight=2.3)

        if load_model_from_directory is not None:
            self.load_weights(
                allowed_characters_for_loaded_model, load_epoch, load_model_from_directory,
                loaded_first_layers_count=frozen_layer_count if reinitialize_trainable_loaded_layers else None)
        if self.data_parallel is not None:
            self.tower.broadcast_parameters()

    # ------------------------------------------------------------------ construction
    def create_predictive_net(self) -> PredictiveNet:
        """The 11-layer Conv1D tower of reference net.py:291-341 as a `ConvTower` on one B200."""
        import torch
        from speechless_b200.engine import ConvTower, wav2letter_layers

        if not torch.cuda.is_available():
            raise RuntimeError("speechless_b200.Wav2Letter needs a CUDA device (B200); there is no CPU fallback.")
        device = torch.device(self._device) if self._device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        layers = wav2letter_layers(self.input_size_per_time_step, self.grapheme_encoding.grapheme_set_size,
                                   activation=self.activation, output_activation=self.output_activation,
                                   main_filter_count=self.main_filter_count,
                                   out_filter_count=self.out_filter_count,
                                   use_raw_wave_input=self.use_raw_wave_input)
        # Dropout sits in front of (wave_conv,) striding_conv and inner_conv_1..7, never before the last three
        # layers (never_dropout, net.py:326-330)
        dropout_layers = list(range(len(layers) - 3)) if self.dropout else []
        # `layers[:frozen_layer_count]` of the reference counts the interleaved Dropout layers too
        combined_count = len(layers) + len(dropout_layers)
        if self.frozen_layer_count > 0:
            log("All but {} layers frozen.".format(combined_count - self.frozen_layer_count))
        combined_index, frozen_convs = 0, 0
        for index in range(len(layers)):
            combined_index += 1 if index in dropout_layers else 0
            if combined_index < self.frozen_layer_count:
                frozen_convs += 1
            combined_index += 1
        self.tower = ConvTower(layers, device, PRECISIONS[self.compute_dtype],
                               frozen_layer_count=frozen_convs, dropout=self.dropout, dropout_layers=dropout_layers,
                               dropout_seed=self.seed if self.seed is not None else 0)
        self.tower.init_glorot(self.seed)
        return PredictiveNet(self.tower, self.input_size_per_time_step)

    @property
    def input_to_prediction_length_ratio(self) -> int:
        """Factor by which striding shortens the output (reference net.py:343-348)."""
        return reduce(lambda x, y: x * y, [layer.strides[0] for layer in self.predictive_net.conv_layers], 1)

    # ------------------------------------------------------------------ weight loading (net.py:184-269)
    @staticmethod
    def indices_to_load_by_target_index(allowed_characters_for_loaded_model: List[chr],
                                        allowed_characters: List[chr]) -> List[Optional[int]]:
        load_character_set = set(allowed_characters_for_loaded_model)
        target_character_set = set(allowed_characters)
        ignored = load_character_set - target_character_set
        if ignored:
            log("Ignoring characters {} from loaded model.".format(sorted(ignored)))
        extra = target_character_set - load_character_set
        if extra:
            log("Initializing extra characters {} not found in model.".format(sorted(extra)))

        character_mapping = [
            single_or_none([index for index, character in enumerate(allowed_characters_for_loaded_model)
                            if character == target_character])
            for target_character in allowed_characters]
        log("Character mapping: {}".format(character_mapping))
        return character_mapping

    def load_weights(self, allowed_characters_for_loaded_model: List[chr], load_epoch: int,
                     load_model_from_directory: Path, loaded_first_layers_count: Optional[int] = None):
        if allowed_characters_for_loaded_model is None:
            self.predictive_net.load_weights(str(Path(load_model_from_directory) / self.model_file_name(load_epoch)))
            return

        layer_count = len(self.predictive_net.layers)
        if loaded_first_layers_count is None:
            loaded_first_layers_count = layer_count

        original_wav2letter = Wav2Letter(input_size_per_time_step=self.input_size_per_time_step,
                                         allowed_characters=allowed_characters_for_loaded_model,
                                         use_raw_wave_input=self.use_raw_wave_input,
                                         activation=self.activation,
                                         output_activation=self.output_activation,
                                         optimizer=self.optimizer,
                                         dropout=self.dropout,
                                         load_model_from_directory=load_model_from_directory,
                                         load_epoch=load_epoch,
                                         frozen_layer_count=self.frozen_layer_count,
                                         use_asg=self.use_asg,
                                         asg_initial_probabilities=self.asg_initial_probabilities,
                                         asg_transition_probabilities=self.asg_transition_probabilities,
                                         main_filter_count=self.main_filter_count,
                                         out_filter_count=self.out_filter_count,
                                         compute_dtype=self.compute_dtype, device=self._device)

        log("Loading first {} layers of {}, epoch {}, reinitializing the last {}.".format(
            loaded_first_layers_count, load_model_from_directory, load_epoch,
            layer_count - loaded_first_layers_count))

        for index, layer in enumerate(self.predictive_net.layers[:loaded_first_layers_count]):
            if isinstance(layer, _DropoutLayerView):
                continue  # (the reference would fail to unpack a Dropout layer's empty weight list here)
            original_weights, original_biases = original_wav2letter.predictive_net.layers[index].get_weights()

            if index == layer_count - 1:
                # re-index the grapheme axis of output_conv for the new alphabet (net.py:240-267)
                mapping = self.indices_to_load_by_target_index(allowed_characters_for_loaded_model,
                                                               self.grapheme_encoding.allowed_characters)
                blank = self.grapheme_encoding.ctc_blank

                def source_index(target_grapheme_index: int) -> Optional[int]:
                    if target_grapheme_index == blank:
                        return original_wav2letter.grapheme_encoding.ctc_blank
                    return mapping[target_grapheme_index]

                sources = [source_index(g) for g in range(self.grapheme_encoding.grapheme_set_size)]
                k, cin, _ = original_weights.shape
                # NB the reference tests `if index`, so source index 0 (its first character) is
                # treated like a missing character and zero-initialised (net.py:254,258); kept.
                new_weights = zeros((k, cin, len(sources)), dtype=original_weights.dtype)
                new_biases = zeros((len(sources),), dtype=original_biases.dtype)
                for target, source in enumerate(sources):
                    if source:
                        new_weights[:, :, target] = original_weights[:, :, source]
                        new_biases[target] = original_biases[source]
                original_weights, original_biases = new_weights, new_biases

            layer.set_weights([original_weights, original_biases])

    # ------------------------------------------------------------------ inference (net.py:350-357, 485-498)
    def prediction_batch(self, input_batch: ndarray) -> ndarray:
        """Grapheme probabilities (B, ceil(T/ratio), V) for a zero-padded spectrogram batch (B, T, F)."""
        ws = self.tower.upload(input_batch)
        self.tower.forward(ws)
        return ws.probs.cpu().numpy()

    def logits_batch(self, input_batch: ndarray) -> ndarray:
        """Pre-softmax activations of output_conv — not part of the reference surface; parity tests use it."""
        ws = self.tower.upload(input_batch)
        self.tower.forward(ws, want_logits=True)
        return ws.logits.cpu().numpy()

    def predict_batch_greedily(self, spectrograms: List[ndarray]) -> List[str]:
        input_batch, prediction_lengths = self._input_batch_and_prediction_lengths(spectrograms)
        return self.grapheme_encoding.decode_prediction_batch(self.prediction_batch(input_batch),
                                                              prediction_lengths=prediction_lengths)

    def get_predicted_graphemes_and_loss_batch(self, input_by_name: Dict[str, ndarray]) -> Tuple[ndarray, ndarray]:
        """One fused device pass: greedy-decoded graphemes (dense, -1 padded, as tf.sparse_to_dense
        in net.py:436) and the per-example CTC loss (B, 1) (net.py:456-459)."""
        if self.use_asg:
            raise NotImplementedError("ASG is not yet implemented.")
        names = Wav2Letter.InputNames
        tower = self.tower
        ws = tower.upload(input_by_name[names.input_batch])
        tower.forward(ws)
        tower.set_labels(ws, input_by_name[names.label_batch], input_by_name[names.prediction_lengths],
                         input_by_name[names.label_lengths])
        loss = tower.ctc(ws, want_grad=False)
        if self.kenlm_directory is not None:
            return self._beam_search_with_language_model(ws), loss.cpu().numpy().reshape(-1, 1)
        decoded, decoded_lengths = tower.greedy_decode(ws, merge_repeated=True)
        decoded = decoded.cpu().numpy()
        width = max(int(decoded_lengths.max().item()), 1)
        return decoded[:, :width], loss.cpu().numpy().reshape(-1, 1)

    def beam_search_batch(self, ws, beam_width: Optional[int] = None, top_paths: int = 1,
                          merge_repeated: bool = False, language_model=None) -> List[List[Tuple[List[int], float]]]:
        """Device prefix beam search over a forwarded workspace: per utterance the `top_paths` best
        (grapheme indices, log-probability) pairs, best first (tf.nn.ctc_beam_search_decoder semantics; the
        reference calls it with merge_repeated=False, net.py:441-447)."""
        decoded, lengths, log_probabilities = self.tower.beam_search_decode(
            ws, beam_width=beam_width or self.decoder_beam_width, top_paths=top_paths, merge_repeated=merge_repeated,
            language_model=language_model)
        decoded, lengths, log_probabilities = decoded.cpu().numpy(), lengths.cpu().numpy(), log_probabilities.cpu().numpy()
        return [[(decoded[b, p, :lengths[b, p]].tolist(), float(log_probabilities[b, p]))
                 for p in range(decoded.shape[1]) if numpy.isfinite(log_probabilities[b, p])]
                for b in range(decoded.shape[0])]

    def predict_batch_with_beam_search(self, spectrograms: List[ndarray], beam_width: Optional[int] = None,
                                       top_paths: int = 1, use_language_model: bool = False
                                       ) -> List[List[Tuple[str, float]]]:
        """Beam-search counterpart of `predict_batch_greedily`: per utterance the `top_paths` best
        (text, log-probability) hypotheses of the device prefix beam search; `use_language_model` scores the
        words of `kenlm_directory`'s model inside the search (the scores then include the language-model terms)."""
        if use_language_model and self.device_language_model is None:
            raise ValueError("use_language_model needs kenlm_directory (and language_model_mode='in-search')")
        input_batch, prediction_lengths = self._input_batch_and_prediction_lengths(spectrograms)
        tower = self.tower
        ws = tower.upload(input_batch)
        tower.forward(ws)
        tower.set_prediction_lengths(ws, prediction_lengths)
        return [[(self.grapheme_encoding.decode_graphemes(graphemes, merge_repeated=False), log_probability)
                 for graphemes, log_probability in hypotheses]
                for hypotheses in self.beam_search_batch(
                    ws, beam_width=beam_width, top_paths=top_paths,
                    language_model=self.device_language_model if use_language_model else None)]

    def _beam_search_with_language_model(self, ws) -> ndarray:
        """Dense (B, max length) grapheme matrix, -1 padded like the greedy path: the best hypothesis of the
        device beam search with the language model inside the search, or (`language_model_mode="rescoring"`)
        the hypothesis of the plain search that the language model re-scores best."""
        in_search = self.device_language_model is not None
        n_best = self.beam_search_batch(ws, top_paths=1 if in_search else self.decoder_top_paths, merge_repeated=False,
                                        language_model=self.device_language_model)
        winners = []
        for hypotheses in n_best:
            if in_search:
                winners.append(hypotheses[0][0] if hypotheses else [])
                continue
            if not hypotheses:  # no finite hypothesis (e.g. NaN probabilities): empty transcription, not a crash
                winners.append([])
                continue
            texts = [(self.grapheme_encoding.decode_graphemes(graphemes, merge_repeated=False), log_probability)
                     for graphemes, log_probability in hypotheses]
            best_text, _ = self.rescorer.best(texts)
            winners.append(next(graphemes for (graphemes, _), (text, _) in zip(hypotheses, texts) if text == best_text))
        width = max(max((len(w) for w in winners), default=0), 1)
        dense = -numpy.ones((len(winners), width), dtype=numpy.int32)
        for row, winner in enumerate(winners):
            dense[row, :len(winner)] = winner
        return dense

    def test_and_predict_batch(self, labeled_spectrogram_batch: List[LabeledSpectrogram]) -> ExpectationsVsPredictions:
        input_by_name, dummy_labels = self._inputs_for_loss_net(labeled_spectrogram_batch)
        predicted_graphemes, loss_batch = self.get_predicted_graphemes_and_loss_batch(input_by_name)

        # blank labels are returned as -1 (net.py:467-468)
        predicted_graphemes = predicted_graphemes.copy()
        predicted_graphemes[predicted_graphemes < 0] = self.grapheme_encoding.ctc_blank

        prediction_lengths = list(numpy.squeeze(input_by_name[Wav2Letter.InputNames.prediction_lengths], axis=1))
        losses = list(numpy.squeeze(loss_batch, axis=1))

        # repeats were already merged on the device, so merging is disabled here (net.py:473-475)
        predictions = self.grapheme_encoding.decode_grapheme_batch(predicted_graphemes, prediction_lengths,
                                                                   merge_repeated=False)
        return ExpectationsVsPredictions(
            [ExpectationVsPrediction(predicted=predicted, expected=expected, loss=float(loss))
             for predicted, expected, loss in
             zip(predictions, (e.label for e in labeled_spectrogram_batch), losses)])

    def test_and_predict(self, labeled_spectrogram: LabeledSpectrogram) -> ExpectationVsPrediction:
        # the reference duplicates the example because TF fails on batches of size 1 (net.py:491-495)
        return self.test_and_predict_batch([labeled_spectrogram, labeled_spectrogram]).results[0]

    def predict(self, labeled_spectrogram: LabeledSpectrogram) -> str:
        return self.test_and_predict(labeled_spectrogram).predicted

    def test_and_predict_batch_with_log(self, index: int, batch: List[LabeledSpectrogram]) -> ExpectationsVsPredictions:
        result = self.test_and_predict_batch(batch)
        log(str(result) + " (batch {})".format(index))
        return result

    def test_and_predict_batches(self, labeled_spectrogram_batches: Iterable[List[LabeledSpectrogram]],
                                 data_parallel=None) -> ExpectationsVsPredictionsInBatches:
        """Under data parallelism the batches are dealt round-robin to the ranks (evaluation needs no
        collective on the data path); the per-batch results are gathered on the host and every
        rank returns the complete, ordered result."""
        data_parallel = data_parallel or self.data_parallel
        if data_parallel is None or not data_parallel.active:
            return ExpectationsVsPredictionsInBatches([self.test_and_predict_batch_with_log(index, batch)
                                                       for index, batch in enumerate(labeled_spectrogram_batches)])
        mine = [(index, self.test_and_predict_batch_with_log(index, batch))
                for index, batch in enumerate(labeled_spectrogram_batches)
                if index % data_parallel.world_size == data_parallel.rank]
        everyone = [pair for pairs in data_parallel.gather_objects(mine) for pair in pairs]
        return ExpectationsVsPredictionsInBatches([result for _, result in sorted(everyone, key=lambda p: p[0])])

    def test_and_predict_batches_with_log(
            self, corpus_name: str, batches: Iterable[List[LabeledSpectrogram]]) -> ExpectationsVsPredictionsInBatches:
        result = self.test_and_predict_batches(batches)
        log("{}: {}".format(corpus_name, result))
        return result

    def test_and_predict_grouped_batches(self, grouped_labeled_spectrogram_batches: Dict[str, Iterable[
        List[LabeledSpectrogram]]]) -> ExpectationsVsPredictionsInGroupedBatches:
        return ExpectationsVsPredictionsInGroupedBatches(
            OrderedDict((corpus_name, self.test_and_predict_batches_with_log(corpus_name=corpus_name,
                                                                             batches=labeled_spectrogram_batches))
                        for corpus_name, labeled_spectrogram_batches in grouped_labeled_spectrogram_batches.items()))

    # ------------------------------------------------------------------ training (net.py:541-576)
    def train_on_batch(self, input_by_name: Dict[str, ndarray], global_batch_size: Optional[int] = None,
                       allreduce: Optional[Callable] = None, data_parallel=None) -> float:
        """One optimisation step on this GPU's shard: forward, CTC loss + gradient, backward,
        (gradient all-reduce), Keras-2 Adam.  Objective = mean over the *global* batch (net.py:389).
        Returns the batch-mean loss — the GLOBAL one whenever gradients are all-reduced.  `data_parallel` (a
        `speechless_b200.distributed.DataParallel`, default: the one given to the constructor) pipelines the
        bucketed all-reduce and the per-bucket Adam update with backward (`ConvTower.backward_and_update`);
        `allreduce(grads, loss_sum)` is the plain, non-overlapped hook.  Under data parallelism the caller
        passes this rank's shard, padded to the GLOBAL longest utterance when the single-process result is to
        be reproduced exactly (padding is unmasked, net.py:578-587; `train` does this), and
        `global_batch_size` (or the key of that name in `input_by_name`)."""
        if self.use_asg:
            raise NotImplementedError("ASG is not yet implemented.")
        names = Wav2Letter.InputNames
        tower = self.tower
        ws = tower.upload(input_by_name[names.input_batch], training=True)
        tower.forward(ws, training=True)  # learning phase 1: dropout active (net.py:597-606)
        tower.set_labels(ws, input_by_name[names.label_batch], input_by_name[names.prediction_lengths],
                         input_by_name[names.label_lengths])
        batch_size = input_by_name.get("global_batch_size", global_batch_size if global_batch_size is not None else ws.B)
        loss = tower.ctc(ws, want_grad=True, grad_scale=1.0 / batch_size)
        self.optimizer.iterations += 1
        if allreduce is not None:
            tower.backward(ws)
            loss_sum = loss.sum()
            allreduce(tower.grads, loss_sum)
            tower.adam_step(self.optimizer.lr, self.optimizer.beta_1, self.optimizer.beta_2, self.optimizer.epsilon,
                            self.optimizer.iterations)
        else:
            loss_sum = tower.backward_and_update(ws, loss, self.optimizer,
                                                 data_parallel=data_parallel or self.data_parallel)
        return float(loss_sum.item()) / batch_size

    def fit_batches(self, input_batches: Iterable[Dict[str, ndarray]], global_batch_size: Optional[int] = None,
                    data_parallel=None) -> List[float]:
        """Train on consecutive batches (the dictionaries `_inputs_for_loss_net` builds), pipelined:
        the host->device copy of batch i+1 runs on a copy stream while batch i computes, and the
        per-step loss is read back asynchronously — one host synchronisation at the end.  Returns
        each step's batch-mean loss (the global mean under data parallelism: the loss sums are all-reduced
        with the gradients).  Same arithmetic as calling `train_on_batch` per batch."""
        import torch
        if self.use_asg:
            raise NotImplementedError("ASG is not yet implemented.")
        names = Wav2Letter.InputNames
        tower = self.tower
        data_parallel = data_parallel or self.data_parallel
        iterator = iter(input_batches)
        current = next(iterator, None)
        if current is None:
            return []
        slot = 0
        ws = tower.stage_async(current[names.input_batch], slot)
        device_losses = []
        batch_sizes = []
        while current is not None:
            upcoming = next(iterator, None)
            next_ws = tower.stage_async(upcoming[names.input_batch], slot ^ 1) if upcoming is not None else None
            tower.consume_slot(ws, slot, training=True)
            tower.forward(ws, training=True)
            tower.set_labels(ws, current[names.label_batch], current[names.prediction_lengths],
                             current[names.label_lengths])
            batch_size = current.get("global_batch_size", global_batch_size if global_batch_size is not None else ws.B)
            loss = tower.ctc(ws, want_grad=True, grad_scale=1.0 / batch_size)
            self.optimizer.iterations += 1
            loss_sum = tower.backward_and_update(ws, loss, self.optimizer, data_parallel=data_parallel)
            if len(device_losses) % 64 == 0:  # pinned allocations are slow: one per 64 steps
                host_block = torch.empty((64,), dtype=torch.float32, pin_memory=True)
            host_loss = host_block[len(device_losses) % 64]
            host_loss.copy_(loss_sum, non_blocking=True)  # device -> host read of the step's result
            device_losses.append(host_loss)
            batch_sizes.append(batch_size)
            current, ws, slot = upcoming, next_ws, slot ^ 1
        tower.sync()
        return [float(l) / n for l, n in zip(device_losses, batch_sizes)]

    def train(self,
              labeled_spectrogram_batches: Iterable[List[LabeledSpectrogram]],
              preview_labeled_spectrogram_batch: List[LabeledSpectrogram],
              tensor_board_log_directory: Path,
              net_directory: Path,
              batches_per_epoch: int,
              *,
              epochs: int = 100000000,
              data_parallel=None):
        """Same schedule as the reference's `fit_generator` call (net.py:550-556): log the preview
        batch, then epochs of `batches_per_epoch` steps starting at `load_epoch`, calling the
        callbacks of `create_callbacks` at every epoch end.  `epochs` (keyword-only, default as in
        the reference) bounds the run for tests; a finite batch iterable ends training quietly.

        `data_parallel` (default: the constructor's): every rank is handed the SAME batch iterable — what the
        reference's single process would consume — trains on its contiguous share of each batch, padded to the
        batch's longest utterance, and the gradients of the global-mean objective are all-reduced, so N GPUs
        take exactly the single-GPU optimisation steps.  Rank 0 alone logs and writes checkpoints."""
        data_parallel = data_parallel or self.data_parallel
        if data_parallel is not None and not data_parallel.active:
            data_parallel = None
        leader = data_parallel is None or data_parallel.rank == 0

        def print_preview_batch():
            result = self.test_and_predict_batch(preview_labeled_spectrogram_batch)
            if leader:
                log(result)

        print_preview_batch()
        callbacks = self.create_callbacks(callback=print_preview_batch,
                                          tensor_board_log_directory=tensor_board_log_directory if leader else None,
                                          net_directory=net_directory, save=leader)
        initial_epoch = self.load_epoch if (self.load_epoch is not None) else 0
        batches = _Prefetcher(self._loss_inputs_generator(labeled_spectrogram_batches, data_parallel))
        try:
            for epoch in range(initial_epoch, epochs):
                started = time.time()
                losses = self.fit_batches((inputs for inputs, _dummy in itertools.islice(batches, batches_per_epoch)),
                                          data_parallel=data_parallel)
                if len(losses) < batches_per_epoch:
                    return  # the batch iterable is exhausted
                logs = {"loss": sum(losses) / len(losses), "seconds": time.time() - started}
                for on_epoch_end in callbacks:
                    on_epoch_end(epoch, logs)
        finally:
            batches.close()

    @staticmethod
    def model_file_name(epoch: int) -> str:
        return "weights-epoch{}.h5".format(epoch)

    def create_callbacks(self, callback: Callable[[], None], tensor_board_log_directory: Path, net_directory: Path,
                         callback_step: int = 1, save_step: int = 1, save: bool = True) -> List[Callable]:
        """Epoch-end hooks with the reference's semantics (net.py:562-576): run `callback` every
        `callback_step` epochs, save `weights-epoch{N}` every `save_step` epochs for N > 0.  The
        TensorBoard callback (net.py:574) writes the epoch's scalars as a TensorBoard event file when the
        `tensorboard` package is importable (no TensorFlow needed), and always as a JSON-lines log
        (`scalars.jsonl`) in `tensor_board_log_directory`."""
        event_writer = []  # created on first use, one per training run

        def custom_callback(epoch: int, logs=()):
            if epoch % callback_step == 0:
                callback()
            if save and epoch % save_step == 0 and epoch > 0:
                mkdir(net_directory)
                self.predictive_net.save_weights(str(Path(net_directory) / self.model_file_name(epoch)))

        def scalar_log(epoch: int, logs=()):
            if tensor_board_log_directory is None:
                return
            mkdir(tensor_board_log_directory)
            with (Path(tensor_board_log_directory) / "scalars.jsonl").open("a") as f:
                f.write(json.dumps(dict(epoch=epoch, **dict(logs))) + "\n")
            try:
                from tensorboard.compat.proto.event_pb2 import Event
                from tensorboard.compat.proto.summary_pb2 import Summary
                from tensorboard.summary.writer.event_file_writer import EventFileWriter
            except ImportError:
                return
            if not event_writer:
                event_writer.append(EventFileWriter(str(tensor_board_log_directory)))
            values = [Summary.Value(tag=str(name), simple_value=float(value)) for name, value in dict(logs).items()]
            event_writer[0].add_event(Event(wall_time=time.time(), step=epoch, summary=Summary(value=values)))
            event_writer[0].flush()

        return [scalar_log, custom_callback]

    # ------------------------------------------------------------------ batching (net.py:500-511, 578-607)
    def _loss_inputs_generator(self, labeled_spectrogram_batches: Iterable[List[LabeledSpectrogram]],
                               data_parallel=None) -> Iterable[Tuple[Dict, ndarray]]:
        for labeled_spectrogram_batch in labeled_spectrogram_batches:
            if data_parallel is None:
                yield self._inputs_for_loss_net(labeled_spectrogram_batch)
                continue
            # this rank's contiguous share, padded to the longest utterance of the WHOLE batch: the unmasked
            # zero padding (net.py:578-587) bleeds into the last valid frames, so only the global pad length
            # reproduces the single-process step; labels are padded per shard (the -1 padding is never read)
            shard = data_parallel.shard(labeled_spectrogram_batch)
            longest = max(x.z_normalized_transposed_spectrogram().shape[0] for x in labeled_spectrogram_batch)
            inputs = self._input_dictionary_for_loss_net(shard, pad_to_length=longest)
            inputs["global_batch_size"] = len(labeled_spectrogram_batch)
            yield inputs, zeros((len(shard),))

    def _inputs_for_loss_net(self, labeled_spectrogram_batch: List[LabeledSpectrogram]) -> Tuple[
        Dict[str, ndarray], ndarray]:
        batch_size = len(labeled_spectrogram_batch)
        dummy_labels_for_dummy_loss_function = zeros((batch_size,))
        training_input_dictionary = self._input_dictionary_for_loss_net(
            labeled_spectrogram_batch=labeled_spectrogram_batch)
        return training_input_dictionary, dummy_labels_for_dummy_loss_function

    def _input_batch_and_prediction_lengths(self, spectrograms: List[ndarray],
                                            pad_to_length: Optional[int] = None) -> Tuple[ndarray, List[int]]:
        """Zero-pad to the longest utterance; prediction length = T // ratio (floor, net.py:582)
        although the tower emits ceil(T / ratio) frames.  Padded frames are NOT masked."""
        batch_size = len(spectrograms)
        input_size_per_time_step = spectrograms[0].shape[1]
        input_lengths = [spectrogram.shape[0] for spectrogram in spectrograms]
        prediction_lengths = [s // self.input_to_prediction_length_ratio for s in input_lengths]
        padded_length = max(input_lengths) if pad_to_length is None else max(pad_to_length, max(input_lengths))
        input_batch = zeros((batch_size, padded_length, input_size_per_time_step), dtype=numpy.float32)
        for index, spectrogram in enumerate(spectrograms):
            input_batch[index, :spectrogram.shape[0], :spectrogram.shape[1]] = spectrogram
        return input_batch, prediction_lengths

    def _prediction_length_batch(self, prediction_lengths: List[int], batch_size: int) -> ndarray:
        return reshape(array(prediction_lengths), (batch_size, 1))

    def _input_dictionary_for_loss_net(self, labeled_spectrogram_batch: List[LabeledSpectrogram],
                                       pad_to_length: Optional[int] = None) -> Dict[str, ndarray]:
        spectrograms = [x.z_normalized_transposed_spectrogram() for x in labeled_spectrogram_batch]
        labels = [x.label for x in labeled_spectrogram_batch]
        input_batch, prediction_lengths = self._input_batch_and_prediction_lengths(spectrograms, pad_to_length)
        label_lengths = reshape(array([len(label) for label in labels]), (len(labeled_spectrogram_batch), 1))
        return {
            Wav2Letter.InputNames.input_batch: input_batch,
            Wav2Letter.InputNames.prediction_lengths: self._prediction_length_batch(prediction_lengths,
                                                                                    batch_size=len(spectrograms)),
            Wav2Letter.InputNames.label_batch: self.grapheme_encoding.encode_label_batch(labels),
            Wav2Letter.InputNames.label_lengths: label_lengths,
            'keras_learning_phase': array([True])
        }
