"""Host-side mirror of the reference's prover interface, calling the CUDA library.

Reference interfaces mirrored (paths relative to the reference tree):
  SRS.new / record fields      src/Sonic/SRS.hs:11-43
  commitPoly, openPoly         src/Sonic/CommitmentScheme.hs:20-48
  prove, Proof, RndOracle      src/Sonic/Protocol.hs:28-109
  hscProve, HscProof           src/Sonic/Signature.hs:22-72
  ArithCircuit, Assignment, GateWeights   bulletproofs records used at Protocol.hs:17

Values: Fr is a Python int in [0, r); a G1 element is its 48-byte compressed encoding
(`bytes`); a `VLaurent Fr` is a dict {exponent: coefficient}.  Nothing here computes on
the CPU beyond packing bytes: all field, curve and polynomial arithmetic happens on the GPU.
"""
from __future__ import annotations

import ctypes
from ctypes import c_uint64, c_void_p
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

from . import capi
from .capi import FAMILY_ALPHA, FAMILY_PLAIN, SonicError, check, lib

R_MODULUS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
G1Bytes = bytes
Laurent = Dict[int, int]


def _fr(x: int) -> bytes:
    return (x % R_MODULUS).to_bytes(32, "little")


def _frs(xs: Sequence[int]) -> bytes:
    return b"".join((x % R_MODULUS).to_bytes(32, "little") for x in xs)


def _dense(f: Laurent) -> Tuple[int, int, bytes]:
    """Sparse Laurent polynomial -> (lo, len, coefficient bytes); zero coefficients dropped
    first, as the reference's normal form does."""
    nz = {e: c % R_MODULUS for e, c in f.items() if c % R_MODULUS}
    if not nz:
        return 0, 0, b""
    lo, hi = min(nz), max(nz)
    zero = bytes(32)
    return lo, hi - lo + 1, b"".join(_fr(nz[e]) if e in nz else zero for e in range(lo, hi + 1))


@dataclass
class GateWeights:
    wL: List[List[int]]
    wR: List[List[int]]
    wO: List[List[int]]


@dataclass
class ArithCircuit:
    weights: GateWeights
    cs: List[int]
    sparse: bool = False  # hand the weights over in CSR form (zero weights dropped) instead of dense Q x n
    _handle: object = field(default=None, repr=False, compare=False)

    def handle(self) -> "_Circuit":
        if self._handle is None:
            self._handle = _Circuit(self)
        return self._handle


@dataclass
class Assignment:
    aL: List[int]
    aR: List[int]
    aO: List[int]


class _Circuit:
    """Weights resident on the device across proofs (sonic_circuit_load)."""

    def __init__(self, c: ArithCircuit):
        w = c.weights
        if not w.wL or not w.wL[0]:
            raise SonicError(1, "Empty weights")
        self.Q, self.n = len(w.wL), len(w.wL[0])
        # the C side copies Q*n (and Q) field elements from these buffers: every shape is checked here.
        # (The reference indexes rows with `!!` and would fail on a ragged matrix too.)
        for name, m in (("wL", w.wL), ("wR", w.wR), ("wO", w.wO)):
            if len(m) != self.Q:
                raise SonicError(1, f"{name} has {len(m)} rows, expected Q = {self.Q}")
            for q, row in enumerate(m):
                if len(row) != self.n:
                    raise SonicError(1, f"{name}[{q}] has {len(row)} entries, expected n = {self.n}")
        if len(c.cs) != self.Q:
            raise SonicError(1, f"cs has {len(c.cs)} entries, expected Q = {self.Q}")
        capi.init()
        h = c_void_p()
        if c.sparse:
            import numpy as np

            args, keep = [], []
            for m in (w.wL, w.wR, w.wO):
                rowptr, cols, vals = [0], [], []
                for row in m:
                    for i, v in enumerate(row):
                        if v % R_MODULUS:
                            cols.append(i)
                            vals.append(v)
                    rowptr.append(len(cols))
                rp = np.array(rowptr, dtype=np.uint64)
                cl = np.array(cols, dtype=np.uint32)
                vl = np.frombuffer(_frs(vals), dtype=np.uint8).copy() if vals else np.zeros(0, dtype=np.uint8)
                keep += [rp, cl, vl]
                args += [rp.ctypes.data, cl.ctypes.data if len(cols) else None, vl.ctypes.data if len(cols) else None]
            check(lib().sonic_circuit_load_csr(self.n, self.Q, *args, _frs(c.cs), ctypes.byref(h)))
        else:
            flat = lambda m: _frs([v for row in m for v in row])
            check(lib().sonic_circuit_load(self.n, self.Q, flat(w.wL), flat(w.wR), flat(w.wO), _frs(c.cs), ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().sonic_circuit_free(self.h)
                self.h = None
        except Exception:
            pass


class SRS:
    """`SRS` record (src/Sonic/SRS.hs:11-22), G1 vectors resident in HBM behind a handle."""

    def __init__(self, handle: c_void_p, d: int):
        self._h = handle
        self.srsD = d

    @staticmethod
    def new(d: int, x: int, alpha: int) -> "SRS":
        """`SRS.new d x alpha` (src/Sonic/SRS.hs:27)."""
        capi.init()
        h = c_void_p()
        check(lib().sonic_srs_new(d, _fr(x), _fr(alpha), ctypes.byref(h)))
        return SRS(h, d)

    def save(self, path: str) -> None:
        """Writes the resident arrays (precomputed levels included) to `path`; no trapdoor is stored."""
        check(lib().sonic_srs_save(self._h, path.encode()))

    @staticmethod
    def load(path: str) -> "SRS":
        capi.init()
        h = c_void_p()
        check(lib().sonic_srs_load(path.encode(), ctypes.byref(h)))
        return SRS(h, int(lib().sonic_srs_d(h)))

    def _range(self, family: int, lo: int, count: int) -> List[G1Bytes]:
        out = ctypes.create_string_buffer(48 * count)
        check(lib().sonic_srs_g1_range(self._h, family, lo, count, out))
        raw = out.raw
        return [raw[48 * i:48 * i + 48] for i in range(count)]

    def g1(self, family: int, exponent: int) -> G1Bytes:
        return self._range(family, exponent, 1)[0]

    # the four G1 record fields, in the reference's index convention
    @property
    def gNegativeX(self) -> List[G1Bytes]:          # [i-1] = g^{x^-i}
        return self._range(FAMILY_PLAIN, -self.srsD, self.srsD)[::-1]

    @property
    def gPositiveX(self) -> List[G1Bytes]:          # [i] = g^{x^i}
        return self._range(FAMILY_PLAIN, 0, self.srsD + 1)

    @property
    def gNegativeAlphaX(self) -> List[G1Bytes]:     # [i-1] = g^{alpha x^-i}
        return self._range(FAMILY_ALPHA, -self.srsD, self.srsD)[::-1]

    @property
    def gPositiveAlphaX(self) -> List[G1Bytes]:     # [i-1] = g^{alpha x^i}
        return self._range(FAMILY_ALPHA, 1, self.srsD)

    def _range_g2(self, family: int, lo: int, count: int) -> List[bytes]:
        out = ctypes.create_string_buffer(96 * count)
        check(lib().sonic_srs_g2_range(self._h, family, lo, count, out))
        raw = out.raw
        return [raw[96 * i:96 * i + 96] for i in range(count)]

    # the four G2 record fields (only with option "g2"); 96-byte compressed G2
    @property
    def hNegativeX(self) -> List[bytes]:
        return self._range_g2(FAMILY_PLAIN, -self.srsD, self.srsD)[::-1]

    @property
    def hPositiveX(self) -> List[bytes]:
        return self._range_g2(FAMILY_PLAIN, 0, self.srsD + 1)

    @property
    def hNegativeAlphaX(self) -> List[bytes]:
        return self._range_g2(FAMILY_ALPHA, -self.srsD, self.srsD)[::-1]

    @property
    def hPositiveAlphaX(self) -> List[bytes]:  # [i] = h^{alpha x^i}, i = 0..d (h^alpha IS shared)
        return self._range_g2(FAMILY_ALPHA, 0, self.srsD + 1)

    def free(self) -> None:
        if self._h:
            lib().sonic_srs_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def commitPoly(srs: SRS, maxm: int, fX: Laurent) -> G1Bytes:
    """`commitPoly srs max f` (src/Sonic/CommitmentScheme.hs:20-33)."""
    lo, ln, coeffs = _dense(fX)
    out = ctypes.create_string_buffer(48)
    check(lib().sonic_commit(srs._h, maxm, lo, ln, coeffs, out))
    return out.raw


def openPoly(srs: SRS, z: int, fX: Laurent) -> Tuple[int, G1Bytes]:
    """`openPoly srs z f` (src/Sonic/CommitmentScheme.hs:36-48) -> (f(z), W)."""
    lo, ln, coeffs = _dense(fX)
    v = ctypes.create_string_buffer(32)
    w = ctypes.create_string_buffer(48)
    check(lib().sonic_open(srs._h, _fr(z), lo, ln, coeffs, v, w))
    return int.from_bytes(v.raw, "little"), w.raw


def msm(srs: SRS, family: int, lo: int, scalars) -> G1Bytes:
    """sum_i scalars[i] * base[family][lo+i]: the fold of CommitmentScheme.hs:26-29 on its own.
    `scalars`: list of ints, or a bytes / numpy buffer of 32-byte little-endian values."""
    if isinstance(scalars, (list, tuple)):
        n, data = len(scalars), _frs(scalars)
    else:
        data = scalars
        n = (len(scalars) if isinstance(scalars, (bytes, bytearray)) else scalars.nbytes) // 32
    out = ctypes.create_string_buffer(48)
    check(lib().sonic_msm_g1(srs._h, family, lo, n, capi.buf(data) if not isinstance(data, bytes) else data, out))
    return out.raw


def msm_partial(srs: SRS, family: int, lo: int, scalars) -> bytes:
    """One slice of a sharded MSM -> 96 raw bytes (see `g1_sum`)."""
    if isinstance(scalars, (list, tuple)):
        n, data = len(scalars), _frs(scalars)
    else:
        data = scalars
        n = (len(scalars) if isinstance(scalars, (bytes, bytearray)) else scalars.nbytes) // 32
    out = ctypes.create_string_buffer(96)
    check(lib().sonic_msm_g1_partial(srs._h, family, lo, n, capi.buf(data) if not isinstance(data, bytes) else data, out))
    return out.raw


def g1_sum(raw_partials: Sequence[bytes]) -> G1Bytes:
    """`<>` over the partial sums gathered from the ranks."""
    capi.init()
    out = ctypes.create_string_buffer(48)
    data = b"".join(raw_partials)
    check(lib().sonic_g1_sum(data, len(raw_partials), out))
    return out.raw


@dataclass
class HscProof:
    """src/Sonic/Signature.hs:22-29."""
    hscS: List[Tuple[G1Bytes, Tuple[int, G1Bytes]]]
    hscW: List[Tuple[int, G1Bytes, G1Bytes]]
    hscQv: G1Bytes
    hscC: G1Bytes
    hscU: int
    hscV: int


@dataclass
class Proof:
    """src/Sonic/Protocol.hs:28-38."""
    prR: G1Bytes
    prT: G1Bytes
    prA: int
    prWa: G1Bytes
    prB: int
    prWb: G1Bytes
    prWt: G1Bytes
    prS: int
    prHscProof: HscProof


@dataclass
class RndOracle:
    """src/Sonic/Protocol.hs:41-45."""
    rndOracleY: int
    rndOracleZ: int
    rndOracleYZs: List[Tuple[int, int]]


def _parse_hsc(buf: bytes, m: int, pos: int = 0) -> HscProof:
    def G():
        nonlocal pos
        v = buf[pos:pos + 48]
        pos += 48
        return v

    def F():
        nonlocal pos
        v = int.from_bytes(buf[pos:pos + 32], "little")
        pos += 32
        return v

    hscS = []
    for _ in range(m):
        cm = G(); s = F(); w = G()
        hscS.append((cm, (s, w)))
    hscW = []
    for _ in range(m):
        sp = F(); wp = G(); qj = G()
        hscW.append((sp, wp, qj))
    qv = G(); c = G(); u = F(); v = F()
    return HscProof(hscS, hscW, qv, c, u, v)


def parse_proof(buf: bytes, Q: int) -> Proof:
    G = lambda o: buf[o:o + 48]
    F = lambda o: int.from_bytes(buf[o:o + 32], "little")
    return Proof(prR=G(0), prT=G(48), prA=F(96), prWa=G(128), prB=F(176), prWb=G(208), prWt=G(256),
                 prS=F(304), prHscProof=_parse_hsc(buf, Q, 336))


def _check_prove_inputs(ch: "_Circuit", assignment: Assignment, rnd: Sequence[int]) -> None:
    """The C ABI reads n Fr from each of aL/aR/aO and 2Q+8 Fr from rnd: lengths are checked before any pointer crosses."""
    if not (len(assignment.aL) == len(assignment.aR) == len(assignment.aO) == ch.n):
        raise SonicError(1, "assignment length differs from the circuit's n")
    if len(rnd) != 2 * ch.Q + 8:
        raise SonicError(1, "prove draws 2Q+8 random field elements")


def prove_bytes(srs: SRS, assignment: Assignment, circuit: ArithCircuit, rnd: Sequence[int]) -> bytes:
    """The boundary call itself: one `sonic_prove`, proof bytes in record order."""
    ch = circuit.handle()
    n, Q = ch.n, ch.Q
    _check_prove_inputs(ch, assignment, rnd)
    size = int(lib().sonic_proof_size(Q))
    out = ctypes.create_string_buffer(size)
    written = c_uint64(0)
    check(lib().sonic_prove(srs._h, ch.h, _frs(assignment.aL), _frs(assignment.aR), _frs(assignment.aO),
                            _frs(rnd), out, size, ctypes.byref(written)))
    return out.raw[:written.value]


def prove(srs: SRS, assignment: Assignment, circuit: ArithCircuit, rnd: Sequence[int]) -> Tuple[Proof, RndOracle]:
    """`prove srs assignment circuit` (src/Sonic/Protocol.hs:47-109).  `rnd` supplies the
    MonadRandom draws in the reference's order: c_{n+1..n+4}, y, z, ys[Q], zs[Q], u, v."""
    buf = prove_bytes(srs, assignment, circuit, rnd)
    Q = circuit.handle().Q
    rnd = [r % R_MODULUS for r in rnd]
    oracle = RndOracle(rnd[4], rnd[5], list(zip(rnd[6:6 + Q], rnd[6 + Q:6 + 2 * Q])))
    return parse_proof(buf, Q), oracle


def hscProve(srs: SRS, circuit: ArithCircuit, yzs: Sequence[Tuple[int, int]], u: int, v: int) -> HscProof:
    """`hscProve srs sXY yzs` (src/Sonic/Signature.hs:32-72); s(X,Y) is the circuit's
    (`sPoly weights`), `u` and `v` are its two `rnd` draws."""
    ch = circuit.handle()
    m = len(yzs)
    if any(len(p) != 2 for p in yzs):
        raise SonicError(1, "yzs holds (y_j, z_j) pairs")
    size = (4 * m + 2) * 48 + (2 * m + 2) * 32
    out = ctypes.create_string_buffer(size)
    written = c_uint64(0)
    flat = _frs([x for pair in yzs for x in pair])
    check(lib().sonic_hsc_prove(srs._h, ch.h, m, flat, _frs([u, v]), out, size, ctypes.byref(written)))
    return _parse_hsc(out.raw, m)


def hscProveBiV(srs: SRS, sXY: Dict[int, Dict[int, int]], yzs: Sequence[Tuple[int, int]], u: int, v: int) -> HscProof:
    """`hscProve srs sXY yzs` for ANY sparse `BiVLaurent Fr` (src/Sonic/Signature.hs:32-37; outer variable
    X, inner Y: {eX: {eY: coeff}}), through sonic_hsc_prove_terms -- what the reference's own test calls
    with `sPoly weights` (test/Test/Signature.hs:30-36)."""
    import numpy as np

    capi.init()
    terms = [(ex, ey, c % R_MODULUS) for ex, inner in sXY.items() for ey, c in inner.items()]
    m = len(yzs)
    if any(len(p) != 2 for p in yzs):
        raise SonicError(1, "yzs holds (y_j, z_j) pairs")
    eX = np.array([t[0] for t in terms], dtype=np.int64)
    eY = np.array([t[1] for t in terms], dtype=np.int64)
    coeff = _frs([t[2] for t in terms])
    size = (4 * m + 2) * 48 + (2 * m + 2) * 32
    out = ctypes.create_string_buffer(size)
    written = c_uint64(0)
    flat = _frs([x for pair in yzs for x in pair])
    check(lib().sonic_hsc_prove_terms(srs._h, len(terms), eX.ctypes.data if terms else None, eY.ctypes.data if terms else None,
                                      coeff if terms else None, m, flat, _frs([u, v]), out, size, ctypes.byref(written)))
    return _parse_hsc(out.raw, m)


def prove_shard(srs: SRS, assignment: Assignment, circuit: ArithCircuit, rnd: Sequence[int], rank: int, world: int) -> bytes:
    """This rank's share of one proof (sonic_prove_shard): raw partial sums + field values."""
    ch = circuit.handle()
    _check_prove_inputs(ch, assignment, rnd)
    if not (world >= 2 and 0 <= rank < world):
        raise SonicError(1, "need world >= 2 and 0 <= rank < world")
    size = int(lib().sonic_shard_blob_size(ch.Q))
    out = ctypes.create_string_buffer(size)
    written = c_uint64(0)
    check(lib().sonic_prove_shard(srs._h, ch.h, _frs(assignment.aL), _frs(assignment.aR), _frs(assignment.aO),
                                  _frs(rnd), rank, world, out, size, ctypes.byref(written)))
    return out.raw[:written.value]


def prove_batch(srs: SRS, assignments: Sequence[Assignment], circuit: ArithCircuit, rnds: Sequence[Sequence[int]]) -> List[bytes]:
    """`mapM (prove srs ?? circuit)` over independent assignments in one call (sonic_prove_batch): with
    several devices whole proofs are dealt round-robin and run at the same time."""
    ch = circuit.handle()
    if len(assignments) != len(rnds):
        raise SonicError(1, "one list of draws per assignment")
    for a, r in zip(assignments, rnds):
        _check_prove_inputs(ch, a, r)
    count = len(assignments)
    size = int(lib().sonic_proof_size(ch.Q))
    out = ctypes.create_string_buffer(size * max(count, 1))
    written = c_uint64(0)
    flat_a = b"".join(_frs(a.aL) + _frs(a.aR) + _frs(a.aO) for a in assignments)
    flat_r = b"".join(_frs(r) for r in rnds)
    check(lib().sonic_prove_batch(srs._h, ch.h, count, flat_a, flat_r, out, size * count, ctypes.byref(written)))
    raw = out.raw
    return [raw[i * size:(i + 1) * size] for i in range(count)]


def prove_combine(Q: int, blobs: Sequence[bytes]) -> bytes:
    """Folds the gathered shard blobs into the proof bytes (sonic_prove_combine)."""
    capi.init()
    want = int(lib().sonic_shard_blob_size(Q))
    if not blobs or any(len(b) != want for b in blobs):
        raise SonicError(1, f"every shard blob must be {want} bytes")
    size = int(lib().sonic_proof_size(Q))
    out = ctypes.create_string_buffer(size)
    written = c_uint64(0)
    check(lib().sonic_prove_combine(Q, len(blobs), b"".join(blobs), out, size, ctypes.byref(written)))
    return out.raw[:written.value]


def pcv_fold(checks: Sequence[Tuple[G1Bytes, int, Tuple[int, G1Bytes], int]], weights: Sequence[int]):
    """G1 side of a batch of `pcV` checks (src/Sonic/CommitmentScheme.hs:51-68): `checks` holds
    (F, z, (v, W), group) with `group` numbering the distinct `max` values; returns (A, B, [C_m])
    such that all checks hold iff e(A, h^{alpha x}) e(B, h^alpha) == prod_m e(C_m, h^{x^{-d+max_m}})
    (up to the soundness error of the random `weights`).  The pairings stay on the host."""
    import numpy as np

    capi.init()
    k = len(checks)
    ng = max(c[3] for c in checks) + 1
    F = b"".join(c[0] for c in checks)
    W = b"".join(c[2][1] for c in checks)
    grp = np.array([c[3] for c in checks], dtype=np.uint32)
    out = ctypes.create_string_buffer(48 * (2 + ng))
    check(lib().sonic_pcv_fold(k, F, W, _frs([c[2][0] for c in checks]), _frs([c[1] for c in checks]), _frs(weights),
                               grp.ctypes.data, ng, out))
    raw = out.raw
    return raw[0:48], raw[48:96], [raw[96 + 48 * m:144 + 48 * m] for m in range(ng)]
