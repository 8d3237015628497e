"""ctypes binding of libspeechless_b200.so (the C-ABI declared in include/speechless_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
product path raises.  PyTorch tensors are used only as device storage (`.data_ptr()`).
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p
from pathlib import Path

PREC_BF16 = 1
PREC_BF16X2 = 2
PREC_FP16 = 3
PRECISIONS = {"bf16": PREC_BF16, "bf16x2": PREC_BF16X2, "fp16": PREC_FP16}
ACT_NONE, ACT_RELU, ACT_SOFTMAX = 0, 1, 2

_LIB_NAME = "libspeechless_b200.so"
_lib = None

# name -> (restype, argtypes); must list every symbol of include/speechless_b200.h
SIGNATURES = {
    "sl_version": (c_int, []),
    "sl_last_error": (c_int, [c_char_p, c_size_t]),
    "sl_sync_check": (c_int, []),
    "sl_pack_activation": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sl_window_activation": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                     ctypes.c_uint64, c_void_p]),
    "sl_unpack_activation": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sl_pack_weights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sl_conv1d_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sl_conv1d_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "sl_dropout_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                               ctypes.c_uint64, c_void_p]),
    "sl_conv1d_dgrad_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "sl_conv1d_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_float, c_void_p]),
    "sl_weights_keras_to_internal": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sl_weights_internal_to_keras": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sl_pack_weights_internal": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sl_ctc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sl_ctc_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "sl_ctc_greedy_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p]),
    "sl_ctc_beam_search_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sl_ctc_beam_search_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                          c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "sl_ctc_beam_search_decode_lm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                             c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                             c_void_p]),
    "sl_spectrogram": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_void_p]),
    "sl_z_normalize": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "sl_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float, c_float, c_float,
                             c_int, c_void_p]),
    "sl_adam_step_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int, c_float, c_float, c_float, c_float, c_int, c_void_p]),
    "sl_comm_unique_id": (c_int, [c_void_p]),
    "sl_comm_init_rank": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_int, c_int, c_int]),
    "sl_comm_size": (c_int, [c_void_p]),
    "sl_allreduce_sum": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "sl_comm_destroy": (c_int, [c_void_p]),
    "sl_comm_nccl_version": (c_int, []),
    "sl_set_sm_limit": (c_int, [c_int]),
}
COMM_ID_BYTES = 128


def library_path() -> Path:
    override = os.environ.get("SPEECHLESS_B200_LIB")
    return Path(override) if override else Path(__file__).resolve().parent / _LIB_NAME


def _try_build(path: Path) -> None:
    """The built library normally travels with the tree; on a fresh checkout compile it once with
    nvcc (sm_100a) if the toolchain is there.  Failure is not fatal here: load() raises."""
    import shutil
    import subprocess
    csrc = path.parent / "csrc"
    if not (csrc / "Makefile").exists() or shutil.which("make") is None:
        return
    if shutil.which("nvcc") is None and not Path("/usr/local/cuda/bin/nvcc").exists():
        return
    try:
        subprocess.run(["make", "-C", str(csrc), "-j", str(os.cpu_count() or 4), "all"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
    except Exception:
        pass


def load():
    """Load (once) and return the ctypes handle; raises RuntimeError when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists() and "SPEECHLESS_B200_LIB" not in os.environ:
        _try_build(path)
    if not path.exists():
        raise RuntimeError(
            "{} not found at {} — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C speechless_b200/csrc`. There is no CPU fallback.".format(_LIB_NAME, path))
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(2048)
    load().sl_last_error(buf, len(buf))
    return buf.value.decode("utf8", "replace")


def check(rc: int) -> None:
    """Map the C-ABI return code onto the reference's exception types."""
    if rc == 0:
        return
    message = last_error()
    if rc in (1, 3):
        raise ValueError(message)
    raise RuntimeError(message)


def ptr(tensor) -> int:
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if tensor is None else tensor.data_ptr()
