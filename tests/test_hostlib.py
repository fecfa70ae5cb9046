"""CPU: the device arithmetic headers compiled for the host (portable emulation of the PTX
carry chain) against Python big ints: the exact limb algorithms of field.cuh / g1.cuh /
scalar.cuh, without a GPU."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import bls12_381 as bls
from tests.util import from_limbs, limbs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q, R = bls.Q, bls.R


@pytest.fixture(scope="module")
def host():
    src = os.path.join(ROOT, "tests", "hostlib", "hostlib.cpp")
    so = os.path.join(ROOT, "tests", "hostlib", "libsonic_hosttest.so")
    deps = [src] + [os.path.join(ROOT, "sonic_b200", "csrc", f) for f in ("field.cuh", "g1.cuh", "scalar.cuh", "constants.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def _arr(v, n):
    return (ctypes.c_uint32 * n)(*limbs(v, n))


@pytest.mark.parametrize("field", ["fq", "fr"])
def test_montgomery_field_ops(host, field):
    fn, n, p = (host.ht_fq_op, 12, Q) if field == "fq" else (host.ht_fr_op, 8, R)
    Rm = 1 << (32 * n)
    Ri = pow(Rm, -1, p)
    rng = random.Random(31)

    def op(k, a, b=0):
        out = (ctypes.c_uint32 * n)()
        fn(k, _arr(a, n), _arr(b, n), out)
        return from_limbs(out)

    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, Rm % p]
    samples = [(a, b) for a in edge for b in edge] + [(rng.randrange(p), rng.randrange(p)) for _ in range(1500)]
    for a, b in samples:
        assert op(0, a, b) == a * b * Ri % p
        assert op(1, a, b) == (a + b) % p
        assert op(2, a, b) == (a - b) % p
        assert op(3, a) == a * Rm % p
        assert op(4, a) == a * Ri % p
        assert op(6, a) == a * a * Ri % p
        assert op(7, a) == (-a) % p
    for _ in range(10):
        a = rng.randrange(1, p)
        assert op(5, a * Rm % p) == pow(a, -1, p) * Rm % p
    # binary-Euclid inverse (the MSM's final to-affine): edge values and random ones
    assert op(8, 0) == 0
    for a in [1, 2, 3, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, Rm % p, Ri, 1 << 31, 1 << 32, (1 << (p.bit_length() - 1))] + [rng.randrange(1, p) for _ in range(400)]:
        assert op(8, a * Rm % p) == pow(a, -1, p) * Rm % p, a


def test_g1_formulas_all_cases(host):
    Rm = 1 << 384
    rng = random.Random(32)

    def aff(P):
        return _arr(0 if P is None else (P[0] * Rm % Q) | ((P[1] * Rm % Q) << 384), 24)

    def xyzz(P, z):
        if P is None:
            return _arr((Rm % Q) | ((Rm % Q) << 384), 48)
        zz, zzz = z * z % Q, z * z * z % Q
        return _arr((P[0] * zz % Q * Rm % Q) | ((P[1] * zzz % Q * Rm % Q) << 384) | ((zz * Rm % Q) << 768) | ((zzz * Rm % Q) << 1152), 48)

    def to_aff(x):
        out = (ctypes.c_uint32 * 24)()
        host.ht_g1_to_affine(x, out)
        v = from_limbs(out)
        xm, ym = v & ((1 << 384) - 1), v >> 384
        if xm == 0 and ym == 0:
            return None
        Ri = pow(Rm, -1, Q)
        return (xm * Ri % Q, ym * Ri % Q)

    G = bls.G1_GEN
    pts = [None, G, bls.g1_neg(G)] + [bls.g1_mul_gen(rng.randrange(R)) for _ in range(5)]
    for A in pts:
        for B in pts:
            out = (ctypes.c_uint32 * 48)()
            host.ht_g1_op(0, xyzz(A, rng.randrange(2, Q)), aff(B), out)
            assert to_aff(out) == bls.g1_add(A, B)
            host.ht_g1_op(1, xyzz(A, rng.randrange(2, Q)), xyzz(B, rng.randrange(2, Q)), out)
            assert to_aff(out) == bls.g1_add(A, B)
        out = (ctypes.c_uint32 * 48)()
        host.ht_g1_op(2, xyzz(A, rng.randrange(2, Q)), None, out)
        assert to_aff(out) == bls.g1_add(A, A)
        c = (ctypes.c_uint8 * 48)()
        host.ht_g1_compress(aff(A), c)
        assert bytes(c) == bls.g1_compress(A)


def test_signed_digit_recoding(host):
    rng = random.Random(33)
    for c in range(4, 21):
        W = (255 + c - 1) // c
        cases = [0, 1, R - 1, (R - 1) // 2, (R - 1) // 2 + 1, 1 << (c - 1), (1 << (c - 1)) + 1] + [rng.randrange(R) for _ in range(100)]
        for s in cases:
            d = (ctypes.c_int32 * 80)()
            assert host.ht_recode(_arr(s, 8), c, d) == W
            assert sum(int(d[j]) << (c * j) for j in range(W)) % R == s
            assert all(abs(int(d[j])) <= 1 << (c - 1) for j in range(W))


def test_wide_products_and_multiplier_variants(host):
    """wide_mul / wide_sqr / Karatsuba (unreduced 2N-limb products, extreme limb patterns included)
    and every Montgomery multiplier variant, plus the fused a*b +- c*d with a single reduction."""
    rng = random.Random(34)

    def pat(n):
        return sum(rng.choice([0, 0xFFFFFFFF, 0x80000000, rng.getrandbits(32)]) << (32 * i) for i in range(n))

    for K, n in ((4, 4), (6, 6), (8, 8), (12, 12), (112, 12), (108, 8)):
        full = (1 << (32 * n)) - 1
        cases = [(full, full), (full, 1), (0, full), (full - 1, full)] + [(pat(n), pat(n)) for _ in range(1500)]
        for a, b in cases:
            out = (ctypes.c_uint32 * (2 * n))()
            host.ht_wide_mul(K, _arr(a, n), _arr(b, n), out)
            assert from_limbs(out) == a * b, (K, hex(a), hex(b))
    for n in (4, 6, 8, 12):
        full = (1 << (32 * n)) - 1
        for a in [full, 0, 1, full - 1, 1 << (32 * n - 1)] + [pat(n) for _ in range(1500)]:
            out = (ctypes.c_uint32 * (2 * n))()
            host.ht_wide_sqr(n, _arr(a, n), out)
            assert from_limbs(out) == a * a, (n, hex(a))
    RiQ = pow(1 << 384, -1, Q)
    for _ in range(1500):
        a1, b1, a2, b2 = (rng.choice([0, 1, Q - 1, rng.randrange(Q)]) for _ in range(4))
        r1, r2 = (ctypes.c_uint32 * 12)(), (ctypes.c_uint32 * 12)()
        host.ht_fq_mul2(_arr(a1, 12), _arr(b1, 12), _arr(a2, 12), _arr(b2, 12), r1, r2)
        assert (from_limbs(r1), from_limbs(r2)) == (a1 * b1 * RiQ % Q, a2 * b2 * RiQ % Q)
    for fn, n, p in ((host.ht_fq_mulvar, 12, Q), (host.ht_fr_mulvar, 8, R)):
        Ri = pow(1 << (32 * n), -1, p)
        edge = [0, 1, p - 1, (p - 1) // 2]
        samples = [(a, b, c, d) for a in edge for b in edge for c in (0, p - 1) for d in (1, p - 1)]
        samples += [tuple(rng.randrange(p) for _ in range(4)) for _ in range(1500)]
        for a, b, c, d in samples:
            ops = [(0, a * b), (1, a * b), (2, a * b), (3, a * b + c * d), (4, a * b - c * d)]
            if n == 12:
                ops += [(5, a * b), (6, a * b)]  # rolled, 4 and 6 rows per iteration
            for op, want in ops:
                out = (ctypes.c_uint32 * n)()
                fn(op, _arr(a, n), _arr(b, n), _arr(c, n), _arr(d, n), out)
                assert from_limbs(out) == want * Ri % p, (op, n)


def test_g1_decompress_and_scalar_mul(host):
    Rm = 1 << 384
    Ri = pow(Rm, -1, Q)
    rng = random.Random(35)

    def aff_of(buf):
        v = from_limbs(buf)
        xm, ym = v & ((1 << 384) - 1), v >> 384
        return None if xm == 0 and ym == 0 else (xm * Ri % Q, ym * Ri % Q)

    pts = [None, bls.G1_GEN, bls.g1_neg(bls.G1_GEN)] + [bls.g1_mul_gen(rng.randrange(R)) for _ in range(12)]
    for P in pts:
        out = (ctypes.c_uint32 * 24)()
        enc = bls.g1_compress(P)
        assert host.ht_g1_decompress(enc, out) == 1
        assert aff_of(out) == P
    bad = [bytes(48), bytes([0xC0]) + bytes(46) + b"\x01", bytes([0x9F]) + b"\xff" * 47]
    x = 1
    while bls.fq_sqrt((x ** 3 + 4) % Q) is not None:
        x += 1
    bad.append(bytes([0x80 | (x >> 376)]) + (x & ((1 << 376) - 1)).to_bytes(47, "big"))  # x not on the curve
    for enc in bad:
        out = (ctypes.c_uint32 * 24)()
        assert host.ht_g1_decompress(enc, out) == 0
    for P in pts[1:6]:
        for k in (0, 1, 2, R - 1, rng.randrange(R)):
            am = _arr((P[0] * Rm % Q) | ((P[1] * Rm % Q) << 384), 24)
            o = (ctypes.c_uint32 * 48)()
            host.ht_g1_mul_scalar(am, _arr(k, 8), o)
            a = (ctypes.c_uint32 * 24)()
            host.ht_g1_to_affine(o, a)
            assert aff_of(a) == bls.g1_mul(P, k)


def test_g2_arithmetic_and_encoding(host):
    rng = random.Random(36)
    for k in [1, 2, 3, R - 1, 0] + [rng.randrange(R) for _ in range(6)]:
        out = (ctypes.c_uint8 * 96)()
        host.ht_g2_mul_gen(_arr(k, 8), out)
        assert bytes(out) == bls.g2_compress(bls.g2_mul(bls.G2_GEN, k)), k
    # the public ZCash encoding of the G2 generator
    out = (ctypes.c_uint8 * 96)()
    host.ht_g2_mul_gen(_arr(1, 8), out)
    assert bytes(out).hex().startswith("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049")
