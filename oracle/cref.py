"""ctypes binding of the C restatement (oracle/csrc/sonic_ref.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsonic_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "csrc", "sonic_ref.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsonic_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.ref_msm_naive.restype = None
        L.ref_msm_naive.argtypes = [c_void_p, c_void_p, c_uint64, c_int, c_void_p, c_void_p]
        L.ref_msm_pippenger.restype = None
        L.ref_msm_pippenger.argtypes = [c_void_p, c_void_p, c_uint64, c_int, c_int, c_void_p]
        L.ref_srs_new.restype = c_int
        L.ref_srs_new.argtypes = [c_uint64, c_void_p, c_void_p, c_int, c_void_p]
        L.ref_commit.restype = c_int
        L.ref_commit.argtypes = [c_void_p, c_uint64, c_int64, c_int64, c_uint64, c_void_p, c_int, c_void_p, POINTER(c_int64)]
        L.ref_open.restype = c_int
        L.ref_open.argtypes = [c_void_p, c_uint64, c_void_p, c_int64, c_uint64, c_void_p, c_int, c_void_p, c_void_p, POINTER(c_int64)]
        L.ref_prove.restype = c_int
        L.ref_prove.argtypes = [c_void_p, c_uint64, c_uint64, c_uint64] + [c_void_p] * 8 + [c_int, c_void_p, POINTER(c_int64)]
        L.ref_time_fq_mul.restype = c_double
        L.ref_time_fq_mul.argtypes = [c_uint64]
        _lib = L
    return _lib


def _ptr(b):
    if b is None:
        return None
    if isinstance(b, bytes):
        return b
    if hasattr(b, "ctypes"):
        return c_void_p(b.ctypes.data)
    return ctypes.cast(b, c_void_p)


def srs_new(d: int, x: int, alpha: int, threads: int = 1) -> bytes:
    """Raw table [family][k+d], 96 bytes per point (see sonic_ref.c: ref_srs_new)."""
    out = ctypes.create_string_buffer(96 * 2 * (2 * d + 1))
    rc = lib().ref_srs_new(d, x.to_bytes(32, "little"), alpha.to_bytes(32, "little"), threads, out)
    if rc:
        raise ZeroDivisionError("SRS.new: recip 0")
    return out.raw


def msm_naive(points_raw, scalars, n: int, threads: int = 1, raw: bool = False) -> bytes:
    out = ctypes.create_string_buffer(96 if raw else 48)
    if raw:
        lib().ref_msm_naive(_ptr(points_raw), _ptr(scalars), n, threads, None, out)
    else:
        lib().ref_msm_naive(_ptr(points_raw), _ptr(scalars), n, threads, out, None)
    return out.raw


def msm_pippenger(points_raw, scalars, n: int, c: int = 12, threads: int = 1) -> bytes:
    """Bucket-method MSM on the CPU (context baseline, NOT the reference's algorithm)."""
    out = ctypes.create_string_buffer(48)
    lib().ref_msm_pippenger(_ptr(points_raw), _ptr(scalars), n, c, threads, out)
    return out.raw


class RefPanic(Exception):
    def __init__(self, code: int, exponent: int):
        self.code, self.exponent = code, exponent
        super().__init__(f"reference panic code {code} at exponent {exponent}")


def commit(srs_raw: bytes, d: int, maxm: int, lo: int, coeffs: bytes, threads: int = 1) -> bytes:
    out = ctypes.create_string_buffer(48)
    err = c_int64(0)
    rc = lib().ref_commit(srs_raw, d, maxm, lo, len(coeffs) // 32, coeffs, threads, out, ctypes.byref(err))
    if rc:
        raise RefPanic(rc, err.value)
    return out.raw


def open_(srs_raw: bytes, d: int, z: int, lo: int, coeffs: bytes, threads: int = 1):
    v = ctypes.create_string_buffer(32)
    w = ctypes.create_string_buffer(48)
    err = c_int64(0)
    rc = lib().ref_open(srs_raw, d, z.to_bytes(32, "little"), lo, len(coeffs) // 32, coeffs, threads, v, w, ctypes.byref(err))
    if rc:
        raise RefPanic(rc, err.value)
    return int.from_bytes(v.raw, "little"), w.raw


def prove(srs_raw, d: int, n: int, Q: int, wL, wR, wO, cs, aL, aR, aO, rnd, threads: int = 1) -> bytes:
    size = (4 * Q + 7) * 48 + (2 * Q + 5) * 32
    out = ctypes.create_string_buffer(size)
    err = c_int64(0)
    rc = lib().ref_prove(_ptr(srs_raw), d, n, Q, _ptr(wL), _ptr(wR), _ptr(wO), _ptr(cs), _ptr(aL), _ptr(aR), _ptr(aO),
                         _ptr(rnd), threads, out, ctypes.byref(err))
    if rc:
        raise RefPanic(rc, err.value)
    return out.raw


def fq_mul_ns(iters: int = 2_000_000) -> float:
    return float(lib().ref_time_fq_mul(iters))
