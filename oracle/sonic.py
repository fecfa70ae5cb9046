"""CPU restatement of the Sonic prover hot path (sdiehl/sonic), Python big-int edition.

TEST INFRASTRUCTURE ONLY — see the header of ``oracle/bls12_381.py``.
PARITY UNPINNED: the reference's tests are round-trip properties without a
single golden value.  This module restates the
reference line by line and is pinned instead by (1) the trapdoor identities,
(2) the reference's own properties re-run on it, (3) the reference's fixed
circuits and (4) public BLS12-381 vectors.  See tests/test_oracle_*.py.

Two layers:

  * the *literal* layer follows the Haskell modules with sparse (dict) Laurent
    polynomials, bivariate where the reference is bivariate; it is O(n^2) like
    the reference and is meant for n <= ~20;
  * the *dense* layer (``prove_dense``) computes the same proof through
    univariate dense vectors, is cross-checked against the literal layer in
    the tests and is what larger parity cases compare against.

Semantics assumed of ``poly-0.4.0.0`` ``Data.Poly.Sparse.Laurent`` (not on disk;
/root/reference/stack.yaml:5-6): terms sorted by exponent, duplicate exponents
summed, zero coefficients dropped, ``toList`` ascending, ``divide`` exact.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from . import bls12_381 as bls
from .bls12_381 import R, INF, fr_inv, fr_pow


class SonicPanic(Exception):
    """Protolude `panic` of the reference (text kept identical)."""


# ======================================================================================
# Sparse Laurent polynomials: {exponent: coeff}, no zero coefficients stored
# ======================================================================================
Laurent = Dict[int, int]


def l_norm(p: Laurent) -> Laurent:
    return {e: c % R for e, c in p.items() if c % R}


def l_monomial(e: int, c: int) -> Laurent:
    c %= R
    return {e: c} if c else {}


def l_add(a: Laurent, b: Laurent) -> Laurent:
    out = dict(a)
    for e, c in b.items():
        v = (out.get(e, 0) + c) % R
        if v:
            out[e] = v
        else:
            out.pop(e, None)
    return out


def l_neg(a: Laurent) -> Laurent:
    return {e: (-c) % R for e, c in a.items()}


def l_scale(a: Laurent, k: int) -> Laurent:
    k %= R
    return l_norm({e: c * k for e, c in a.items()})


def l_mul(a: Laurent, b: Laurent) -> Laurent:
    out: Dict[int, int] = {}
    for ea, ca in a.items():
        for eb, cb in b.items():
            out[ea + eb] = (out.get(ea + eb, 0) + ca * cb) % R
    return l_norm(out)


def l_eval(a: Laurent, x: int) -> int:
    """`eval` of Data.Poly.Sparse.Laurent: negative powers through `recip`
    (/root/reference/src/Sonic/CommitmentScheme.hs:43)."""
    acc = 0
    for e, c in a.items():
        acc = (acc + c * fr_pow(x, e)) % R
    return acc


def l_to_list(a: Laurent) -> List[Tuple[int, int]]:
    return sorted(a.items())


def l_divide_linear(a: Laurent, z: int) -> Optional[Laurent]:
    """`a `divide` (X - z)` (/root/reference/src/Sonic/CommitmentScheme.hs:44):
    exact division or Nothing."""
    if not a:
        return {}
    if z % R == 0:
        # the divisor [(0,-0),(1,1)] normalises to the monomial X: Laurent division shifts offsets
        return {e - 1: c for e, c in a.items()}
    lo, hi = min(a), max(a)
    # a = X^lo * g(X), g an ordinary polynomial of degree hi-lo
    g = [a.get(lo + k, 0) for k in range(hi - lo + 1)]
    q = [0] * (len(g) - 1)
    carry = 0
    for k in range(len(g) - 1, 0, -1):
        carry = (g[k] + z * carry) % R
        q[k - 1] = carry
    rem = (g[0] + z * carry) % R
    if rem:
        return None
    return l_norm({lo + k: c for k, c in enumerate(q)})


# bivariate: outer variable X, inner variable Y (/root/reference/src/Sonic/Utils.hs:15)
BiV = Dict[int, Laurent]


def bv_norm(p: BiV) -> BiV:
    return {e: c for e, c in p.items() if c}


def bv_add(a: BiV, b: BiV) -> BiV:
    out = dict(a)
    for e, c in b.items():
        v = l_add(out.get(e, {}), c)
        if v:
            out[e] = v
        else:
            out.pop(e, None)
    return out


def bv_mul(a: BiV, b: BiV) -> BiV:
    out: BiV = {}
    for ea, ca in a.items():
        for eb, cb in b.items():
            out[ea + eb] = l_add(out.get(ea + eb, {}), l_mul(ca, cb))
    return bv_norm(out)


def evalX(x: int, p: BiV) -> Laurent:
    """/root/reference/src/Sonic/Utils.hs:17-18 — sum of inner polys scaled by x^e."""
    out: Laurent = {}
    for e, inner in p.items():
        out = l_add(out, l_scale(inner, fr_pow(x, e)))
    return out


def evalY(y: int, p: BiV) -> Laurent:
    """/root/reference/src/Sonic/Utils.hs:20-21 — evaluate every inner poly at y."""
    return l_norm({e: l_eval(inner, y) for e, inner in p.items()})


def fromX(p: Laurent) -> BiV:
    """/root/reference/src/Sonic/Utils.hs:23-24."""
    return {e: l_monomial(0, c) for e, c in p.items()}


def fromY(p: Laurent) -> BiV:
    """/root/reference/src/Sonic/Utils.hs:26-27."""
    return bv_norm({0: dict(p)})


# ======================================================================================
# Circuit types (bulletproofs-1.1.0 records; /root/reference/src/Sonic/Protocol.hs:17)
# ======================================================================================
@dataclass
class GateWeights:
    wL: List[List[int]]
    wR: List[List[int]]
    wO: List[List[int]]


@dataclass
class ArithCircuit:
    weights: GateWeights
    cs: List[int]


@dataclass
class Assignment:
    aL: List[int]
    aR: List[int]
    aO: List[int]


# ======================================================================================
# Sonic.Constraints
# ======================================================================================
def rPoly(a: Assignment) -> BiV:
    """/root/reference/src/Sonic/Constraints.hs:23-31."""
    n = len(a.aL)
    out: BiV = {}
    for i, (ai, bi, ci) in enumerate(zip(a.aL, a.aR, a.aO), start=1):
        for e, c in ((i, ai), (-i, bi), (-i - n, ci)):
            out = bv_add(out, bv_norm({e: l_monomial(e, c)}))
    return out


def sPoly(w: GateWeights) -> BiV:
    """/root/reference/src/Sonic/Constraints.hs:34-53."""
    n = len(w.wL[0])

    def xiY(i: int, xL: List[List[int]]) -> Laurent:
        acc: Laurent = {}
        for q, row in enumerate(xL, start=1):
            acc = l_add(acc, l_monomial(q + n, row[i - 1]))
        return acc

    out: BiV = {}
    for i in range(1, n + 1):
        u = xiY(i, w.wL)
        v = xiY(i, w.wR)
        wi = l_add(l_add(l_monomial(-i, -1), l_monomial(i, -1)), xiY(i, w.wO))
        for e, c in ((-i, u), (i, v), (i + n, wi)):
            out = bv_add(out, bv_norm({e: c}))
    return out


def kPoly(k: Sequence[int], n: int) -> Laurent:
    """/root/reference/src/Sonic/Constraints.hs:67-68."""
    out: Laurent = {}
    for e, c in zip(range(n + 1, n + 1 + len(k)), k):
        out = l_add(out, l_monomial(e, c))
    return out


def tPoly(rXY: BiV, sXY: BiV, kY: Laurent) -> BiV:
    """/root/reference/src/Sonic/Constraints.hs:56-65."""
    rXYp = bv_add(rXY, sXY)
    rX1 = fromX(evalY(1, rXY))
    k1Y = fromY(l_neg(kY))
    return bv_add(bv_mul(rX1, rXYp), k1Y)


# ======================================================================================
# Sonic.SRS
# ======================================================================================
@dataclass
class SRS:
    """G1 half of /root/reference/src/Sonic/SRS.hs:11-22 (G2/GT stay on the host
    library and are out of scope), plus the trapdoor for the oracle's checks."""
    srsD: int
    gNegativeX: List
    gPositiveX: List
    gNegativeAlphaX: List
    gPositiveAlphaX: List
    x: int = 0
    alpha: int = 0


def srs_new(d: int, x: int, alpha: int) -> SRS:
    """/root/reference/src/Sonic/SRS.hs:27-43."""
    xInv = fr_inv(x)
    mg = bls.g1_mul_gen
    return SRS(
        srsD=d,
        gNegativeX=[mg(fr_pow(xInv, i)) for i in range(1, d + 1)],
        gPositiveX=[mg(fr_pow(x, i)) for i in range(0, d + 1)],
        gNegativeAlphaX=[mg(alpha * fr_pow(xInv, i)) for i in range(1, d + 1)],
        gPositiveAlphaX=[mg(alpha * fr_pow(x, i)) for i in range(1, d + 1)],
        x=x % R,
        alpha=alpha % R,
    )


# ======================================================================================
# Sonic.CommitmentScheme
# ======================================================================================
def _index(annot: str, v: List, e: int):
    """/root/reference/src/Sonic/CommitmentScheme.hs:70-73."""
    if 0 <= e < len(v):
        return v[e]
    raise SonicPanic(f"{annot} is not long enough: {e} >= {len(v)}")


def commitPoly(srs: SRS, maxm: int, fX: Laurent):
    """/root/reference/src/Sonic/CommitmentScheme.hs:20-33."""
    difference = srs.srsD - maxm
    xfX = l_to_list(l_mul(l_monomial(difference, 1), fX))
    acc = INF
    for e, v in xfX:
        if e > 0:
            base = _index("commitPoly: gPositiveAlphaX", srs.gPositiveAlphaX, e - 1)
        else:
            base = _index("commitPoly: gNegativeAlphaX", srs.gNegativeAlphaX, abs(e) - 1)
        acc = bls.g1_add(acc, bls.g1_mul(base, v))
    return acc


def openPoly(srs: SRS, z: int, fX: Laurent):
    """/root/reference/src/Sonic/CommitmentScheme.hs:36-48."""
    fz = l_eval(fX, z)
    wPoly = l_divide_linear(l_add(fX, l_monomial(0, -fz)), z)
    if wPoly is None:
        raise SonicPanic("Maybe.fromJust: Nothing")
    acc = INF
    for e, v in l_to_list(wPoly):
        if e >= 0:
            base = _index("openPoly: gPositiveX", srs.gPositiveX, e)
        else:
            base = _index("openPoly: gNegativeX", srs.gNegativeX, abs(e) - 1)
        acc = bls.g1_add(acc, bls.g1_mul(base, v))
    return fz, acc


def pcV_trapdoor(srs: SRS, maxm: int, commitment, z: int, opening) -> bool:
    """`pcV` (/root/reference/src/Sonic/CommitmentScheme.hs:51-68) checked in the
    exponent with the trapdoor instead of three pairings:
        e(W, h^{alpha x}) e(g^v W^{-z}, h^alpha) == e(F, h^{x^{-d+max}})
    <=> alpha*(x*w + v - z*w) == f * x^{-d+max}   for W=g^w, F=g^f.
    Discrete logs are not available, so the check is done on the G1 side:
        alpha*(x - z)*W + alpha*v*G == x^{-d+max} * F.
    """
    v, w = opening
    lhs = bls.g1_add(bls.g1_mul(w, srs.alpha * (srs.x - z)), bls.g1_mul_gen(srs.alpha * v))
    rhs = bls.g1_mul(commitment, fr_pow(srs.x, -srs.srsD + maxm))
    return lhs == rhs


# ======================================================================================
# Sonic.Signature / Sonic.Protocol (prover side)
# ======================================================================================
@dataclass
class HscProof:
    hscS: List[Tuple[object, Tuple[int, object]]]
    hscW: List[Tuple[int, object, object]]
    hscQv: object
    hscC: object
    hscU: int
    hscV: int


@dataclass
class Proof:
    prR: object
    prT: object
    prA: int
    prWa: object
    prB: int
    prWb: object
    prWt: object
    prS: int
    prHscProof: HscProof


def rnd_count(Q: int) -> int:
    """Number of Fr values `prove` draws: 4 blinders, y, z, ys[Q], zs[Q], u, v
    (/root/reference/src/Sonic/Protocol.hs:58,66,76,84-85; Signature.hs:48,60)."""
    return 2 * Q + 8


def hscProve(srs: SRS, sXY: BiV, yzs: Sequence[Tuple[int, int]], u: int, v: int) -> HscProof:
    """/root/reference/src/Sonic/Signature.hs:32-72; `u`, `v` are the two `rnd` draws."""
    ss = []
    for yi, zi in yzs:
        sXy = evalY(yi, sXY)
        cm = commitPoly(srs, srs.srsD, sXy)
        op = openPoly(srs, zi, sXy)
        ss.append((cm, op))
    suX = evalX(u, sXY)
    c = commitPoly(srs, srs.srsD, suX)
    sW = []
    for yi, _zi in yzs:
        _, wjp = openPoly(srs, u, evalY(yi, sXY))
        sjp, qj = openPoly(srs, yi, suX)
        sW.append((sjp, wjp, qj))
    _, qv = openPoly(srs, v, suX)
    return HscProof(hscS=ss, hscW=sW, hscQv=qv, hscC=c, hscU=u % R, hscV=v % R)


def prove(srs: SRS, assignment: Assignment, circuit: ArithCircuit, rnd: Sequence[int]):
    """/root/reference/src/Sonic/Protocol.hs:47-109.  `rnd` supplies the MonadRandom
    draws in the reference's order.  Returns (Proof, (y, z, yzs))."""
    n = len(assignment.aL)
    m = len(circuit.weights.wL)
    if srs.srsD < 7 * n:
        raise SonicPanic(
            f"Parameter d is not large enough: {srs.srsD} should be greater than {7 * n}")
    if len(rnd) != rnd_count(m):
        raise ValueError("wrong number of random draws")
    rnd = [r % R for r in rnd]
    cns = rnd[0:4]
    sumcXY: BiV = {}
    for i, cni in enumerate(cns, start=1):
        e = -(2 * n + i)
        sumcXY = bv_add(sumcXY, bv_norm({e: l_monomial(e, cni)}))
    polyRp = bv_add(rPoly(assignment), sumcXY)
    rX1 = evalY(1, polyRp)
    commitR = commitPoly(srs, n, rX1)
    y = rnd[4]
    kY = kPoly(circuit.cs, n)
    sXY = sPoly(circuit.weights)
    tXY = tPoly(polyRp, sXY, kY)
    tXy = evalY(y, tXY)
    commitT = commitPoly(srs, srs.srsD, tXy)
    z = rnd[5]
    a, wa = openPoly(srs, z, rX1)
    b, wb = openPoly(srs, y * z % R, rX1)
    _, wt = openPoly(srs, z, tXy)
    szy = l_eval(evalY(y, sXY), z)
    ys = rnd[6:6 + m]
    zs = rnd[6 + m:6 + 2 * m]
    yzs = list(zip(ys, zs))
    hsc = hscProve(srs, sXY, yzs, rnd[6 + 2 * m], rnd[7 + 2 * m])
    proof = Proof(prR=commitR, prT=commitT, prA=a, prWa=wa, prB=b, prWb=wb, prWt=wt,
                  prS=szy, prHscProof=hsc)
    return proof, (y, z, yzs)


def hscVerify_trapdoor(srs: SRS, sXY: BiV, yzs, proof: HscProof) -> bool:
    """/root/reference/src/Sonic/Signature.hs:74-90 with `pcV_trapdoor`."""
    sv = l_eval(evalY(proof.hscV, sXY), proof.hscU)
    ok = pcV_trapdoor(srs, srs.srsD, proof.hscC, proof.hscV, (sv, proof.hscQv))
    for (yi, zi), (ci, (si, wi)), (sip, wip, qi) in zip(yzs, proof.hscS, proof.hscW):
        ok = ok and pcV_trapdoor(srs, srs.srsD, ci, zi, (si, wi))
        ok = ok and pcV_trapdoor(srs, srs.srsD, ci, proof.hscU, (sip, wip))
        ok = ok and pcV_trapdoor(srs, srs.srsD, proof.hscC, yi, (sip, qi))
    return ok


def verify_trapdoor(srs: SRS, circuit: ArithCircuit, proof: Proof, y: int, z: int, yzs) -> bool:
    """/root/reference/src/Sonic/Protocol.hs:111-130 with `pcV_trapdoor`."""
    n = len(circuit.weights.wL[0])
    kY = kPoly(circuit.cs, n)
    sXY = sPoly(circuit.weights)
    t = (proof.prA * (proof.prB + proof.prS) - l_eval(kY, y)) % R
    return all([
        hscVerify_trapdoor(srs, sXY, yzs, proof.prHscProof),
        pcV_trapdoor(srs, n, proof.prR, z, (proof.prA, proof.prWa)),
        pcV_trapdoor(srs, n, proof.prR, y * z % R, (proof.prB, proof.prWb)),
        pcV_trapdoor(srs, srs.srsD, proof.prT, z, (t, proof.prWt)),
    ])


# ======================================================================================
# Boundary encoding of a proof: field order of `Proof` / `HscProof`
# ======================================================================================
def proof_size(Q: int) -> int:
    return (4 * Q + 7) * 48 + (2 * Q + 5) * 32


def encode_proof(p: Proof) -> bytes:
    g, f = bls.g1_compress, bls.fr_to_bytes
    out = [g(p.prR), g(p.prT), f(p.prA), g(p.prWa), f(p.prB), g(p.prWb), g(p.prWt), f(p.prS)]
    h = p.prHscProof
    for cm, (s, w) in h.hscS:
        out += [g(cm), f(s), g(w)]
    for sp, wp, qj in h.hscW:
        out += [f(sp), g(wp), g(qj)]
    out += [g(h.hscQv), g(h.hscC), f(h.hscU), f(h.hscV)]
    return b"".join(out)


def decode_proof(buf: bytes, Q: int) -> Proof:
    assert len(buf) == proof_size(Q)
    pos = 0

    def G():
        nonlocal pos
        v = bls.g1_decompress(buf[pos:pos + 48])
        pos += 48
        return v

    def F():
        nonlocal pos
        v = bls.fr_from_bytes(buf[pos:pos + 32])
        pos += 32
        return v

    prR, prT, prA, prWa, prB, prWb, prWt, prS = G(), G(), F(), G(), F(), G(), G(), F()
    hscS = []
    for _ in range(Q):
        cm = G(); s = F(); w = G()
        hscS.append((cm, (s, w)))
    hscW = []
    for _ in range(Q):
        sp = F(); wp = G(); qj = G()
        hscW.append((sp, wp, qj))
    qv, c, u, v = G(), G(), F(), F()
    return Proof(prR, prT, prA, prWa, prB, prWb, prWt, prS,
                 HscProof(hscS, hscW, qv, c, u, v))


# ======================================================================================
# Dense layer: the same values through univariate dense vectors
# ======================================================================================
class Dense:
    """Dense Laurent vector: coefficient of X^(lo+k) is c[k].  Zeros are kept in
    the vector and skipped where the reference's sparse form would not hold them."""

    __slots__ = ("lo", "c")

    def __init__(self, lo: int, c: List[int]):
        self.lo = lo
        self.c = c

    def to_sparse(self) -> Laurent:
        return {self.lo + k: v for k, v in enumerate(self.c) if v}

    def eval(self, x: int) -> int:
        acc = 0
        for v in reversed(self.c):
            acc = (acc * x + v) % R
        return acc * fr_pow(x, self.lo) % R if self.c else 0


def dense_rX1(a: Assignment, cns: Sequence[int]) -> Dense:
    """r'(X,1) over X^{-2n-4..n} (Protocol.hs:58-63)."""
    n = len(a.aL)
    lo = -2 * n - 4
    c = [0] * (3 * n + 5)
    for i in range(1, n + 1):
        c[i - lo] = a.aL[i - 1] % R
        c[-i - lo] = a.aR[i - 1] % R
        c[-i - n - lo] = a.aO[i - 1] % R
    for i in range(1, 5):
        c[-2 * n - i - lo] = cns[i - 1] % R
    return Dense(lo, c)


def dense_sXy(w: GateWeights, y: int) -> Dense:
    """s(X,y) over X^{-n..2n} (Constraints.hs:34-53 then Utils.hs:20-21)."""
    n = len(w.wL[0])
    Qn = len(w.wL)
    lo = -n
    c = [0] * (3 * n + 1)
    ypow = [fr_pow(y, q + n) for q in range(1, Qn + 1)]
    for i in range(1, n + 1):
        c[-i - lo] = sum(ypow[q] * w.wL[q][i - 1] for q in range(Qn)) % R
        c[i - lo] = sum(ypow[q] * w.wR[q][i - 1] for q in range(Qn)) % R
        c[i + n - lo] = (-fr_pow(y, i) - fr_pow(y, -i)
                         + sum(ypow[q] * w.wO[q][i - 1] for q in range(Qn))) % R
    return Dense(lo, c)


def dense_suY(w: GateWeights, u: int) -> Dense:
    """s(u,Y) over Y^{-n..n+Q} (Utils.hs:17-18 applied to Constraints.hs:34-53)."""
    n = len(w.wL[0])
    Qn = len(w.wL)
    lo = -n
    c = [0] * (2 * n + Qn + 1)
    for i in range(1, n + 1):
        t = (-fr_pow(u, i + n)) % R
        c[i - lo] = (c[i - lo] + t) % R
        c[-i - lo] = (c[-i - lo] + t) % R
    for q in range(1, Qn + 1):
        acc = 0
        for i in range(1, n + 1):
            acc += (fr_pow(u, -i) * w.wL[q - 1][i - 1] + fr_pow(u, i) * w.wR[q - 1][i - 1]
                    + fr_pow(u, i + n) * w.wO[q - 1][i - 1])
        c[q + n - lo] = (c[q + n - lo] + acc) % R
    return Dense(lo, c)


def dense_mul(a: Dense, b: Dense) -> Dense:
    out = [0] * (len(a.c) + len(b.c) - 1)
    for i, x in enumerate(a.c):
        if x:
            for j, yv in enumerate(b.c):
                out[i + j] += x * yv
    return Dense(a.lo + b.lo, [v % R for v in out])


def dense_tXy(rX1: Dense, sXy: Dense, y: int, ky: int) -> Dense:
    """t(X,y) = r'(X,1) * (r'(X,y) + s(X,y)) - k(y)."""
    rXy = Dense(rX1.lo, [v * fr_pow(y, rX1.lo + k) % R for k, v in enumerate(rX1.c)])
    lo = min(rXy.lo, sXy.lo)
    hi = max(rXy.lo + len(rXy.c), sXy.lo + len(sXy.c))
    s = [0] * (hi - lo)
    for k, v in enumerate(rXy.c):
        s[rXy.lo + k - lo] = v
    for k, v in enumerate(sXy.c):
        s[sXy.lo + k - lo] = (s[sXy.lo + k - lo] + v) % R
    t = dense_mul(rX1, Dense(lo, s))
    t.c[-t.lo] = (t.c[-t.lo] - ky) % R
    return t


def msm_pippenger(points: Sequence, scalars: Sequence[int], c: int = 8):
    """Bucket-method MSM on big ints (oracle speed-up only; value == naive fold)."""
    pairs = [(p, s % R) for p, s in zip(points, scalars) if p is not INF and s % R]
    if not pairs:
        return INF
    nwin = (255 + c - 1) // c
    total = (1, 1, 0)
    for wdx in reversed(range(nwin)):
        for _ in range(c):
            total = bls._jac_double(total)
        buckets = [None] * ((1 << c) - 1)
        for p, s in pairs:
            dgt = (s >> (wdx * c)) & ((1 << c) - 1)
            if dgt:
                b = buckets[dgt - 1]
                buckets[dgt - 1] = bls._to_jac(p) if b is None else bls._jac_add(b, bls._to_jac(p))
        run = (1, 1, 0)
        acc = (1, 1, 0)
        for b in reversed(buckets):
            if b is not None:
                run = bls._jac_add(run, b)
            acc = bls._jac_add(acc, run)
        total = bls._jac_add(total, acc)
    return bls._from_jac(total)


def srs_base(srs: SRS, alpha_family: bool, k: int):
    """Exponent-indexed view of the SRS: g^{x^k} or g^{alpha x^k}; raises the
    reference's panic text for what `index` would reject."""
    if alpha_family:
        if k > 0:
            return _index("commitPoly: gPositiveAlphaX", srs.gPositiveAlphaX, k - 1)
        return _index("commitPoly: gNegativeAlphaX", srs.gNegativeAlphaX, abs(k) - 1)
    if k >= 0:
        return _index("openPoly: gPositiveX", srs.gPositiveX, k)
    return _index("openPoly: gNegativeX", srs.gNegativeX, abs(k) - 1)


def commit_dense(srs: SRS, maxm: int, f: Dense):
    diff = srs.srsD - maxm
    pts, scs = [], []
    for k, v in enumerate(f.c):
        if v:
            pts.append(srs_base(srs, True, f.lo + k + diff))
            scs.append(v)
    return msm_pippenger(pts, scs)


def open_dense(srs: SRS, z: int, f: Dense):
    z %= R
    if f.lo < 0 and z == 0 and any(f.c):
        raise ZeroDivisionError("Fr: recip 0")
    fz = f.eval(z) if any(f.c) else 0
    g = list(f.c)
    if any(g) or fz:
        # subtract f(z) at X^0 (the window always contains exponent 0 for the callers here)
        if not (f.lo <= 0 < f.lo + len(g)):
            raise ValueError("dense window must contain X^0")
        g[-f.lo] = (g[-f.lo] - fz) % R
    q = [0] * max(len(g) - 1, 0)
    carry = 0
    for k in range(len(g) - 1, 0, -1):
        carry = (g[k] + z * carry) % R
        q[k - 1] = carry
    assert not g or (g[0] + z * carry) % R == 0
    pts, scs = [], []
    for k, v in enumerate(q):
        if v:
            pts.append(srs_base(srs, False, f.lo + k))
            scs.append(v)
    return fz, msm_pippenger(pts, scs)


def prove_dense(srs: SRS, assignment: Assignment, circuit: ArithCircuit, rnd: Sequence[int]):
    """Same outputs as `prove`, through dense univariate vectors."""
    n = len(assignment.aL)
    m = len(circuit.weights.wL)
    if srs.srsD < 7 * n:
        raise SonicPanic(
            f"Parameter d is not large enough: {srs.srsD} should be greater than {7 * n}")
    rnd = [r % R for r in rnd]
    assert len(rnd) == rnd_count(m)
    w = circuit.weights
    rX1 = dense_rX1(assignment, rnd[0:4])
    commitR = commit_dense(srs, n, rX1)
    y, z = rnd[4], rnd[5]
    ky = sum(k * fr_pow(y, n + 1 + q) for q, k in enumerate(circuit.cs)) % R
    sXy = dense_sXy(w, y)
    tXy = dense_tXy(rX1, sXy, y, ky)
    commitT = commit_dense(srs, srs.srsD, tXy)
    a, wa = open_dense(srs, z, rX1)
    b, wb = open_dense(srs, y * z % R, rX1)
    _, wt = open_dense(srs, z, tXy)
    szy = sXy.eval(z)
    ys = rnd[6:6 + m]
    zs = rnd[6 + m:6 + 2 * m]
    u, v = rnd[6 + 2 * m], rnd[7 + 2 * m]
    sXyj = [dense_sXy(w, yj) for yj in ys]
    ss = []
    for sj, zj in zip(sXyj, zs):
        ss.append((commit_dense(srs, srs.srsD, sj), open_dense(srs, zj, sj)))
    suY = dense_suY(w, u)
    c = commit_dense(srs, srs.srsD, suY)
    sW = []
    for sj, yj in zip(sXyj, ys):
        _, wjp = open_dense(srs, u, sj)
        sjp, qj = open_dense(srs, yj, suY)
        sW.append((sjp, wjp, qj))
    _, qv = open_dense(srs, v, suY)
    proof = Proof(commitR, commitT, a, wa, b, wb, wt, szy,
                  HscProof(ss, sW, qv, c, u, v))
    return proof, (y, z, list(zip(ys, zs)))


# ======================================================================================
# The proof as a list of MSMs (record order) -- used to emulate sharded proving in tests
# ======================================================================================
def prove_dense_plan(srs: SRS, assignment: Assignment, circuit: ArithCircuit, rnd: Sequence[int]):
    """Returns (msms, fvals): msms = [(alpha_family, lo_exponent, scalars)] in the order the
    proof record lists its G1 fields, fvals = the Fr fields in record order.  Folding every
    MSM gives exactly `prove_dense`."""
    n = len(assignment.aL)
    m = len(circuit.weights.wL)
    d = srs.srsD
    rnd = [r % R for r in rnd]
    w = circuit.weights

    def quotient(f: Dense, z: int):
        fz = f.eval(z)
        g = list(f.c)
        g[-f.lo] = (g[-f.lo] - fz) % R
        q = [0] * (len(g) - 1)
        carry = 0
        for k in range(len(g) - 1, 0, -1):
            carry = (g[k] + z * carry) % R
            q[k - 1] = carry
        return fz, (False, f.lo, q)

    def commit(f: Dense, maxm: int):
        return (True, f.lo + d - maxm, list(f.c))

    rX1 = dense_rX1(assignment, rnd[0:4])
    y, z = rnd[4], rnd[5]
    ky = sum(k * fr_pow(y, n + 1 + q) for q, k in enumerate(circuit.cs)) % R
    sXy = dense_sXy(w, y)
    tXy = dense_tXy(rX1, sXy, y, ky)
    ys, zs = rnd[6:6 + m], rnd[6 + m:6 + 2 * m]
    u, v = rnd[6 + 2 * m], rnd[7 + 2 * m]
    a, qa = quotient(rX1, z)
    b, qb = quotient(rX1, y * z % R)
    _, qt = quotient(tXy, z)
    msms = [commit(rX1, n), commit(tXy, d), qa, qb, qt]
    fvals = [a, b, sXy.eval(z)]
    sj = [dense_sXy(w, yj) for yj in ys]
    suY = dense_suY(w, u)
    s_vals, sp_vals = [], []
    for f, zj in zip(sj, zs):
        val, q = quotient(f, zj)
        msms += [commit(f, d), q]
        s_vals.append(val)
    for f, yj in zip(sj, ys):
        _, q1 = quotient(f, u)
        val, q2 = quotient(suY, yj)
        msms += [q1, q2]
        sp_vals.append(val)
    _, qv = quotient(suY, v)
    msms += [qv, commit(suY, d)]
    fvals += s_vals + sp_vals + [u, v]
    return msms, fvals


def fold_msm(srs: SRS, msm, lo_clip: Optional[int] = None, hi_clip: Optional[int] = None):
    """Sum of one planned MSM, optionally restricted to exponents in [lo_clip, hi_clip)."""
    alpha_family, lo, scal = msm
    pts, scs = [], []
    for k, v in enumerate(scal):
        e = lo + k
        if lo_clip is not None and not (lo_clip <= e < hi_clip):
            continue
        if v:
            pts.append(srs_base(srs, alpha_family, e))
            scs.append(v)
    return msm_pippenger(pts, scs)


def assemble_proof_bytes(Q: int, g48: Sequence[bytes], fvals: Sequence[int]) -> bytes:
    """Record-order interleaving of the 4Q+7 G1 encodings and 2Q+5 Fr values."""
    g, f = list(g48), [bls.fr_to_bytes(v) for v in fvals]
    out = [g[0], g[1], f[0], g[2], f[1], g[3], g[4], f[2]]
    gi, fi = 5, 3
    for _ in range(Q):
        out += [g[gi], f[fi], g[gi + 1]]
        gi += 2
        fi += 1
    for _ in range(Q):
        out += [f[fi], g[gi], g[gi + 1]]
        gi += 2
        fi += 1
    out += [g[gi], g[gi + 1], f[fi], f[fi + 1]]
    return b"".join(out)


def verify_trapdoor_dense(srs_d: int, x: int, alpha: int, circuit: ArithCircuit, proof: Proof, y: int, z: int, yzs) -> bool:
    """`verify` (/root/reference/src/Sonic/Protocol.hs:111-130 with Signature.hs:74-90) through
    dense vectors and the trapdoor form of `pcV`; needs no SRS vectors, so it scales to n = 2^16."""
    n = len(circuit.weights.wL[0])
    w = circuit.weights
    srs = SRS(srsD=srs_d, gNegativeX=[], gPositiveX=[], gNegativeAlphaX=[], gPositiveAlphaX=[], x=x % R, alpha=alpha % R)
    ky = sum(k * fr_pow(y, n + 1 + q) for q, k in enumerate(circuit.cs)) % R
    t = (proof.prA * (proof.prB + proof.prS) - ky) % R
    h = proof.prHscProof
    ok = pcV_trapdoor(srs, n, proof.prR, z, (proof.prA, proof.prWa))
    ok = ok and pcV_trapdoor(srs, n, proof.prR, y * z % R, (proof.prB, proof.prWb))
    ok = ok and pcV_trapdoor(srs, srs_d, proof.prT, z, (t, proof.prWt))
    sv = dense_sXy(w, h.hscV).eval(h.hscU)
    ok = ok and pcV_trapdoor(srs, srs_d, h.hscC, h.hscV, (sv, h.hscQv))
    for (yi, zi), (ci, (si, wi)), (sip, wip, qi) in zip(yzs, h.hscS, h.hscW):
        ok = ok and pcV_trapdoor(srs, srs_d, ci, zi, (si, wi))
        ok = ok and pcV_trapdoor(srs, srs_d, ci, h.hscU, (sip, wip))
        ok = ok and pcV_trapdoor(srs, srs_d, h.hscC, yi, (sip, qi))
    return bool(ok)
