"""ctypes binding of include/sonic_b200.h (the same symbols a `foreign import ccall` binds)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsonic_b200.so")

SONIC_OK = 0
ERR_NAMES = {
    1: "INVALID_ARG", 2: "SRS_TOO_SHORT", 3: "D_TOO_SMALL", 4: "DIV_BY_ZERO", 5: "NONCANONICAL",
    6: "CUDA", 7: "BUFFER_TOO_SMALL", 8: "NO_DEVICE", 9: "NOT_INITIALISED",
}
FAMILY_PLAIN, FAMILY_ALPHA = 0, 1


class SonicError(RuntimeError):
    """A non-zero status from the C ABI; `.text` is what the reference passes to `panic`."""

    def __init__(self, code: int, text: str):
        self.code = code
        self.kind = ERR_NAMES.get(code, str(code))
        self.text = text
        super().__init__(f"[{self.kind}] {text}")


_lib = None
_ready = False

# every symbol include/sonic_b200.h declares: name -> (restype, argtypes)
_u8p = c_void_p
SYMBOLS = {
    "sonic_init": (c_int, [POINTER(c_int), c_int]),
    "sonic_shutdown": (None, []),
    "sonic_device_count": (c_int, []),
    "sonic_strerror": (c_char_p, [c_int]),
    "sonic_last_error": (c_size_t, [ctypes.c_char_p, c_size_t]),
    "sonic_srs_new": (c_int, [c_uint64, _u8p, _u8p, POINTER(c_void_p)]),
    "sonic_srs_free": (None, [c_void_p]),
    "sonic_srs_d": (c_uint64, [c_void_p]),
    "sonic_srs_save": (c_int, [c_void_p, c_char_p]),
    "sonic_srs_load": (c_int, [c_char_p, POINTER(c_void_p)]),
    "sonic_srs_g1": (c_int, [c_void_p, c_int, c_int64, _u8p]),
    "sonic_srs_g1_range": (c_int, [c_void_p, c_int, c_int64, c_uint64, _u8p]),
    "sonic_srs_g2_range": (c_int, [c_void_p, c_int, c_int64, c_uint64, _u8p]),
    "sonic_commit": (c_int, [c_void_p, c_int64, c_int64, c_uint64, _u8p, _u8p]),
    "sonic_open": (c_int, [c_void_p, _u8p, c_int64, c_uint64, _u8p, _u8p, _u8p]),
    "sonic_msm_g1": (c_int, [c_void_p, c_int, c_int64, c_uint64, _u8p, _u8p]),
    "sonic_msm_g1_partial": (c_int, [c_void_p, c_int, c_int64, c_uint64, _u8p, _u8p]),
    "sonic_g1_sum": (c_int, [_u8p, c_uint64, _u8p]),
    "sonic_msm_g1_device": (c_int, [c_void_p, c_int, c_int64, c_uint64, c_void_p, _u8p]),
    "sonic_msm_g1_device_partial": (c_int, [c_void_p, c_int, c_int64, c_uint64, c_void_p, _u8p]),
    "sonic_circuit_load": (c_int, [c_uint64, c_uint64, _u8p, _u8p, _u8p, _u8p, POINTER(c_void_p)]),
    "sonic_circuit_load_csr": (c_int, [c_uint64, c_uint64] + [c_void_p] * 10 + [POINTER(c_void_p)]),
    "sonic_circuit_free": (None, [c_void_p]),
    "sonic_rnd_count": (c_uint64, [c_uint64]),
    "sonic_proof_size": (c_uint64, [c_uint64]),
    "sonic_prove": (c_int, [c_void_p, c_void_p, _u8p, _u8p, _u8p, _u8p, _u8p, c_uint64, POINTER(c_uint64)]),
    "sonic_prove_batch": (c_int, [c_void_p, c_void_p, c_uint64, _u8p, _u8p, _u8p, c_uint64, POINTER(c_uint64)]),
    "sonic_shard_exchange_size": (c_uint64, [c_uint64]),
    "sonic_shard_blob_size": (c_uint64, [c_uint64]),
    "sonic_prove_shard": (c_int, [c_void_p, c_void_p, _u8p, _u8p, _u8p, _u8p, ctypes.c_uint32, ctypes.c_uint32, _u8p, c_uint64, POINTER(c_uint64)]),
    "sonic_prove_shard_sink": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, _u8p, ctypes.c_uint32, ctypes.c_uint32, _u8p, c_uint64, POINTER(c_uint64), c_void_p]),
    "sonic_prove_combine_device": (c_int, [c_uint64, ctypes.c_uint32, c_void_p, _u8p, _u8p, c_uint64, POINTER(c_uint64)]),
    "sonic_prove_combine": (c_int, [c_uint64, ctypes.c_uint32, _u8p, _u8p, c_uint64, POINTER(c_uint64)]),
    "sonic_prove_device": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _u8p, _u8p, c_uint64, POINTER(c_uint64)]),
    "sonic_prove_shard_device": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _u8p, ctypes.c_uint32, ctypes.c_uint32, _u8p, c_uint64, POINTER(c_uint64)]),
    "sonic_hsc_prove": (c_int, [c_void_p, c_void_p, c_uint64, _u8p, _u8p, _u8p, c_uint64, POINTER(c_uint64)]),
    "sonic_hsc_prove_terms": (c_int, [c_void_p, c_uint64, c_void_p, c_void_p, _u8p, c_uint64, _u8p, _u8p, _u8p, c_uint64, POINTER(c_uint64)]),
    "sonic_pcv_fold": (c_int, [c_uint64, _u8p, _u8p, _u8p, _u8p, _u8p, c_void_p, ctypes.c_uint32, _u8p]),
    "sonic_set_option": (c_int, [c_char_p, c_int64]),
    "sonic_last_timing_ms": (c_double, [c_char_p]),
    "sonic_last_timing_ms_dev": (c_double, [c_int, c_char_p]),
    "sonic_launch_count": (c_uint64, []),
    "sonic_bench_mark": (c_int, [c_int]),
    "sonic_bench_elapsed_ms": (c_double, [c_int, c_int]),
    "sonic_imad_peak_lmacs": (c_double, [c_int, c_int]),
    "sonic_selftest_field": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, ctypes.c_uint32]),
    "sonic_selftest_g1": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_uint32]),
    "sonic_selftest_latency_ns": (c_double, [c_int, c_int, c_int, c_int]),
    "sonic_dev_alloc": (c_int, [c_uint64, POINTER(c_void_p)]),
    "sonic_dev_free": (c_int, [c_void_p]),
    "sonic_dev_upload": (c_int, [c_void_p, c_void_p, c_uint64]),
    "sonic_dev_download": (c_int, [c_void_p, c_void_p, c_uint64]),
}


def _preload_nccl() -> None:
    """libsonic_b200.so links libnccl.so.2 (the in-library multi-GPU exchange).  A Python process that
    also imports torch must end up with ONE NCCL: torch's wheel bundles its own under the same soname,
    and the loader keeps whichever came first.  Load the bundled one first when it is installed, so that
    the order of `import torch` / `import sonic_b200` does not matter; a process without torch gets the
    system library through the normal search path."""
    if os.environ.get("SONIC_NO_NCCL_PRELOAD"):
        return
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            path = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(path):
                ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    except Exception:
        pass


def lib():
    """Loads libsonic_b200.so; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SonicError(8, f"{LIB_PATH} is missing: build it with `make` / `__graft_entry__.build()`; "
                                "there is no CPU fallback")
        _preload_nccl()
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(1024)
    lib().sonic_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != SONIC_OK:
        raise SonicError(rc, last_error())


def init(device=None) -> None:
    """Binds this process to one GPU (an int; default LOCAL_RANK under torchrun, else device 0) or to
    several GPUs of the box (a list of CUDA ordinals: the library then shards SRS.new, prove and
    prove_batch over them itself, include/sonic_b200.h: sonic_init)."""
    global _ready
    if _ready:
        return
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    devices = list(device) if isinstance(device, (list, tuple)) else [int(device)]
    dev = (c_int * len(devices))(*devices)
    check(lib().sonic_init(dev, len(devices)))
    _ready = True
    import atexit

    atexit.register(shutdown)   # NCCL communicators and worker threads are torn down before the interpreter goes


def device_count() -> int:
    return int(lib().sonic_device_count())


def shutdown() -> None:
    global _ready
    if _ready:
        lib().sonic_shutdown()
        _ready = False


def set_option(name: str, value: int) -> None:
    check(lib().sonic_set_option(name.encode(), value))


def last_timing_ms(stage: str = "total", slot: int = 0) -> float:
    return float(lib().sonic_last_timing_ms_dev(slot, stage.encode()))


def launch_count() -> int:
    return int(lib().sonic_launch_count())


def buf(b) -> c_void_p:
    """Pointer to a bytes-like / numpy / ctypes buffer without copying."""
    if b is None:
        return c_void_p(0)
    if isinstance(b, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(b)) if isinstance(b, bytes) else (ctypes.c_char * len(b)).from_buffer(b), c_void_p)
    if hasattr(b, "ctypes"):  # numpy
        return c_void_p(b.ctypes.data)
    if hasattr(b, "data_ptr"):  # torch
        return c_void_p(b.data_ptr())
    return ctypes.cast(b, c_void_p)
