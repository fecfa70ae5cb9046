"""GPU: the field and curve primitives, through the C ABI self-test hooks, against the
big-int oracle.  Integer work: the bar is bit-exact."""
import ctypes
import random

import numpy as np
import pytest

from oracle import bls12_381 as bls
from tests.util import from_limbs, limbs

pytestmark = pytest.mark.gpu
Q, R = bls.Q, bls.R


def _run_field(sb, field, op, a_vals, b_vals):
    nl = 12 if field == 0 else 8
    n = len(a_vals)
    a = np.array([limbs(v, nl) for v in a_vals], dtype=np.uint32)
    b = np.array([limbs(v, nl) for v in b_vals], dtype=np.uint32)
    out = np.zeros_like(a)
    from sonic_b200 import capi
    capi.check(capi.lib().sonic_selftest_field(field, op, a.ctypes.data, b.ctypes.data, out.ctypes.data, n))
    return [from_limbs(row) for row in out]


@pytest.mark.parametrize("field", [0, 1])
def test_field_ops_bit_exact(gpu, field):
    p, nl = (Q, 12) if field == 0 else (R, 8)
    Rm = 1 << (32 * nl)
    Ri = pow(Rm, -1, p)
    rng = random.Random(100 + field)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, Rm % p, (1 << (32 * nl - 3)) % p]
    a = edge * len(edge) + [rng.randrange(p) for _ in range(4000)]
    b = [e for e in edge for _ in edge] + [rng.randrange(p) for _ in range(4000)]
    assert _run_field(gpu, field, 0, a, b) == [x * y * Ri % p for x, y in zip(a, b)]
    assert _run_field(gpu, field, 1, a, b) == [(x + y) % p for x, y in zip(a, b)]
    assert _run_field(gpu, field, 2, a, b) == [(x - y) % p for x, y in zip(a, b)]
    assert _run_field(gpu, field, 3, a, b) == [x * Rm % p for x in a]
    assert _run_field(gpu, field, 4, a, b) == [x * Ri % p for x in a]
    assert _run_field(gpu, field, 6, a, b) == [x * x * Ri % p for x in a]
    assert _run_field(gpu, field, 7, a, b) == [(-x) % p for x in a]
    small = a[:64] + a[-64:]
    got = _run_field(gpu, field, 5, small, small)
    want = [0 if x == 0 else pow(x * Ri % p, -1, p) * Rm % p for x in small]
    assert got == want
    # the single-thread inverse used by the MSM's final to-affine (binary Euclid, divergent lanes here)
    small = a[:81] + a[-200:]
    assert _run_field(gpu, field, 8, small, small) == [0 if x == 0 else pow(x * Ri % p, -1, p) * Rm % p for x in small]


def _xyzz(P, z=1):
    Rm = 1 << 384
    if P is None:
        return limbs(Rm % Q, 12) + limbs(Rm % Q, 12) + [0] * 24
    zz = z * z % Q
    zzz = zz * z % Q
    return (limbs(P[0] * zz % Q * Rm % Q, 12) + limbs(P[1] * zzz % Q * Rm % Q, 12)
            + limbs(zz * Rm % Q, 12) + limbs(zzz * Rm % Q, 12))


def _affine_as_xyzz(P):
    Rm = 1 << 384
    if P is None:
        return [0] * 48
    return limbs(P[0] * Rm % Q, 12) + limbs(P[1] * Rm % Q, 12) + [0] * 24


def test_g1_group_law_bit_exact(gpu):
    from sonic_b200 import capi
    rng = random.Random(7)
    G = bls.G1_GEN
    pts = [None, G, bls.g1_neg(G)] + [bls.g1_mul_gen(rng.randrange(R)) for _ in range(9)]
    pairs = [(A, B) for A in pts for B in pts]
    n = len(pairs)
    for op in (0, 1, 2, 3, 4, 5):   # 4, 5: addition / doubling shared by a quad of lanes (csrc/g1coop.cuh)
        za = [rng.randrange(2, Q) for _ in pairs]
        zb = [rng.randrange(2, Q) for _ in pairs]
        if op == 3:
            a = np.array([_affine_as_xyzz(A) for A, _ in pairs], dtype=np.uint32)
        else:
            a = np.array([_xyzz(A, z) for (A, _), z in zip(pairs, za)], dtype=np.uint32)
        if op == 0:
            b = np.array([_affine_as_xyzz(B) for _, B in pairs], dtype=np.uint32)
        else:
            b = np.array([_xyzz(B, z) for (_, B), z in zip(pairs, zb)], dtype=np.uint32)
        aff = np.zeros((n, 24), dtype=np.uint32)
        comp = np.zeros((n, 48), dtype=np.uint8)
        capi.check(capi.lib().sonic_selftest_g1(op, a.ctypes.data, b.ctypes.data, aff.ctypes.data, comp.ctypes.data, n))
        for i, (A, B) in enumerate(pairs):
            want = bls.g1_add(A, B) if op in (0, 1, 4) else bls.g1_add(A, A)
            assert bytes(comp[i]) == bls.g1_compress(want), (op, i)
