"""BLS12-381 Fr / Fq / G1 arithmetic on Python big integers.

TEST INFRASTRUCTURE ONLY.  This file is part of the CPU oracle for the Sonic
prover hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product path (``sonic_b200/``) never does.

PARITY UNPINNED.  The reference (sdiehl/sonic) holds no golden vector for any
G1 point or Fr value, and its arithmetic lives in un-vendored Hackage packages:

  * ``pairing-1.0.0``           (/root/reference/stack.yaml:10)  types Fr, G1
  * ``elliptic-curve-0.3.0``    (/root/reference/stack.yaml:9)   gen, mul, <>
  * ``galois-field-1.0.1@b59ecd8`` (/root/reference/stack.yaml:7-8) pow, recip

What is restated here is the published mathematics those packages implement:
the prime fields Fq and Fr, the short Weierstrass curve E: y^2 = x^3 + 4 over
Fq with its standard generator, the group law, and scalar multiplication.
External anchors: the public BLS12-381 constants and the ZCash compressed
encodings of the identity, G and 2G (checked in tests/test_oracle_pins.py).

Results of the group law are canonical (an affine point or infinity), so any
correct algorithm yields the same value the Haskell packages produce; the
representation used here (Jacobian, windowed multiplication) is free.
"""
from __future__ import annotations

# --- constants (SURVEY.md section 8 header; public BLS12-381 parameters) -------------
Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
G1_X = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
G1_Y = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
CURVE_B = 4
FR_BYTES = 32
FQ_BYTES = 48
FR_TWO_ADICITY = 32
FR_GENERATOR = 7  # multiplicative generator of Fr*

INF = None  # the point at infinity (`mempty` of the G1 monoid)
G1_GEN = (G1_X, G1_Y)


# --- Fr -------------------------------------------------------------------------------
def fr(x: int) -> int:
    return x % R


def fr_inv(x: int) -> int:
    """`recip` of galois-field's Prime field; 0 has no inverse (the Haskell
    library throws `divide by zero`)."""
    x %= R
    if x == 0:
        raise ZeroDivisionError("Fr: recip 0")
    return pow(x, R - 2, R)


def fr_pow(x: int, e: int) -> int:
    """`pow` of Data.Field.Galois: negative exponents go through `recip`
    (/root/reference/src/Sonic/SRS.hs:33-41 uses both signs)."""
    if e >= 0:
        return pow(x % R, e, R)
    return pow(fr_inv(x), -e, R)


def fr_to_bytes(x: int) -> bytes:
    """Boundary encoding: 32-byte little-endian canonical residue."""
    return (x % R).to_bytes(FR_BYTES, "little")


def fr_from_bytes(b: bytes) -> int:
    v = int.from_bytes(b, "little")
    if v >= R:
        raise ValueError("Fr encoding is not canonical")
    return v


# --- G1: affine in, affine out; Jacobian inside ---------------------------------------
def g1_is_on_curve(p) -> bool:
    if p is INF:
        return True
    x, y = p
    return (y * y - (x * x * x + CURVE_B)) % Q == 0


def _to_jac(p):
    if p is INF:
        return (1, 1, 0)
    return (p[0], p[1], 1)


def _from_jac(j):
    X, Y, Z = j
    if Z == 0:
        return INF
    zi = pow(Z, Q - 2, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 * zi % Q)


def _jac_double(j):
    X, Y, Z = j
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    A = X * X % Q
    B = Y * Y % Q
    C = B * B % Q
    D = 2 * ((X + B) * (X + B) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def _jac_add(j1, j2):
    X1, Y1, Z1 = j1
    X2, Y2, Z2 = j2
    if Z1 == 0:
        return j2
    if Z2 == 0:
        return j1
    Z1Z1 = Z1 * Z1 % Q
    Z2Z2 = Z2 * Z2 % Q
    U1 = X1 * Z2Z2 % Q
    U2 = X2 * Z1Z1 % Q
    S1 = Y1 * Z2 * Z2Z2 % Q
    S2 = Y2 * Z1 * Z1Z1 % Q
    if U1 == U2:
        if S1 == S2:
            return _jac_double(j1)
        return (1, 1, 0)
    H = (U2 - U1) % Q
    I = 4 * H * H % Q
    J = H * I % Q
    r = 2 * (S2 - S1) % Q
    V = U1 * I % Q
    X3 = (r * r - J - 2 * V) % Q
    Y3 = (r * (V - X3) - 2 * S1 * J) % Q
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % Q
    return (X3, Y3, Z3)


def g1_add(p, q):
    """`<>` on G1 (/root/reference/src/Sonic/CommitmentScheme.hs:26,45)."""
    return _from_jac(_jac_add(_to_jac(p), _to_jac(q)))


def g1_neg(p):
    if p is INF:
        return INF
    return (p[0], (-p[1]) % Q)


def g1_mul(p, k: int):
    """`mul` of Data.Curve: scalar multiplication by an Fr element
    (/root/reference/src/Sonic/CommitmentScheme.hs:27-28,46-47, SRS.hs:33-41)."""
    k %= R
    if k == 0 or p is INF:
        return INF
    acc = (1, 1, 0)
    base = _to_jac(p)
    for bit in bin(k)[2:]:
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add(acc, base)
    return _from_jac(acc)


def g1_sum(points):
    acc = (1, 1, 0)
    for p in points:
        acc = _jac_add(acc, _to_jac(p))
    return _from_jac(acc)


class FixedBase:
    """Windowed fixed-base multiplication of one point (speeds the oracle's
    SRS generation up; the values are those of `g1_mul`)."""

    def __init__(self, p, window: int = 8):
        self.w = window
        self.nwin = (255 + window - 1) // window
        self.table = []
        base = _to_jac(p)
        for _ in range(self.nwin):
            row = [(1, 1, 0)]
            for _ in range((1 << window) - 1):
                row.append(_jac_add(row[-1], base))
            # keep the row affine-normalised so additions below stay cheap
            self.table.append(row)
            for _ in range(window):
                base = _jac_double(base)

    def mul_jac(self, k: int):
        k %= R
        acc = (1, 1, 0)
        mask = (1 << self.w) - 1
        i = 0
        while k:
            dgt = k & mask
            if dgt:
                acc = _jac_add(acc, self.table[i][dgt])
            k >>= self.w
            i += 1
        return acc

    def mul(self, k: int):
        return _from_jac(self.mul_jac(k))


_GEN_TABLE = None


def g1_mul_gen(k: int):
    """`mul gen k` (/root/reference/src/Sonic/SRS.hs:33-41)."""
    global _GEN_TABLE
    if _GEN_TABLE is None:
        _GEN_TABLE = FixedBase(G1_GEN, 8)
    return _GEN_TABLE.mul(k)


def g1_msm_naive(points, scalars):
    """The reference's commitment loop: one scalar multiplication per term,
    folded with point additions (/root/reference/src/Sonic/CommitmentScheme.hs:26-29)."""
    acc = (1, 1, 0)
    for p, s in zip(points, scalars):
        if p is INF or s % R == 0:
            continue
        acc = _jac_add(acc, _to_jac(g1_mul(p, s)))
    return _from_jac(acc)


# --- boundary encoding of G1 --------------------------
# The reference never serialises a point; the 48-byte ZCash-style compressed
# form named in BASELINE.json's north_star is a bijection of the affine value.
def g1_compress(p) -> bytes:
    if p is INF:
        return bytes([0xC0]) + bytes(47)
    x, y = p
    b = bytearray(x.to_bytes(FQ_BYTES, "big"))
    b[0] |= 0x80
    if y > (Q - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def fq_sqrt(a: int):
    # q = 3 (mod 4)
    s = pow(a, (Q + 1) // 4, Q)
    return s if s * s % Q == a % Q else None


def g1_decompress(b: bytes):
    if len(b) != FQ_BYTES or not (b[0] & 0x80):
        raise ValueError("not a compressed G1 encoding")
    if b[0] & 0x40:
        if (b[0] & 0x3F) or any(b[1:]):
            raise ValueError("bad infinity encoding")
        return INF
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= Q:
        raise ValueError("x not canonical")
    y = fq_sqrt((x * x * x + CURVE_B) % Q)
    if y is None:
        raise ValueError("x is not on the curve")
    if (y > (Q - 1) // 2) != bool(b[0] & 0x20):
        y = Q - y
    return (x, y)


def g1_to_raw(p) -> bytes:
    """Uncompressed little-endian (x||y) 96-byte form used for device-table dumps; infinity = zeros."""
    if p is INF:
        return bytes(96)
    return p[0].to_bytes(48, "little") + p[1].to_bytes(48, "little")


def g1_from_raw(b: bytes):
    x = int.from_bytes(b[:48], "little")
    y = int.from_bytes(b[48:96], "little")
    if x == 0 and y == 0:
        return INF
    return (x, y)


# ======================================================================================
# G2 over Fq2 = Fq[u]/(u^2 + 1): the h-vectors of the SRS (/root/reference/src/Sonic/SRS.hs:35-36,40-41)
# Out of the prover's path (only pcV reads four of them); restated for the G2 fixed-base batch
# that SURVEY.md section 8f lists as a "next" item.  E': y^2 = x^3 + 4(1 + u).
# ======================================================================================
G2_X = (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
        0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e)
G2_Y = (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
        0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)
G2_GEN = (G2_X, G2_Y)
G2_B = (4, 4)


def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_sqr(a):
    return f2_mul(a, a)


def f2_scale(a, k):
    return (a[0] * k % Q, a[1] * k % Q)


def f2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % Q, Q - 2, Q)
    return (a[0] * n % Q, (-a[1]) * n % Q)


F2_ZERO, F2_ONE = (0, 0), (1, 0)


def g2_is_on_curve(p) -> bool:
    if p is INF:
        return True
    x, y = p
    return f2_sqr(y) == f2_add(f2_mul(f2_sqr(x), x), G2_B)


def _g2_jac_double(j):
    X, Y, Z = j
    if Z == F2_ZERO:
        return (F2_ONE, F2_ONE, F2_ZERO)
    A = f2_sqr(X)
    B = f2_sqr(Y)
    C = f2_sqr(B)
    t = f2_add(X, B)
    D = f2_scale(f2_sub(f2_sub(f2_sqr(t), A), C), 2)
    E = f2_scale(A, 3)
    F = f2_sqr(E)
    X3 = f2_sub(F, f2_scale(D, 2))
    Y3 = f2_sub(f2_mul(E, f2_sub(D, X3)), f2_scale(C, 8))
    Z3 = f2_scale(f2_mul(Y, Z), 2)
    return (X3, Y3, Z3)


def _g2_jac_add(j1, j2):
    X1, Y1, Z1 = j1
    X2, Y2, Z2 = j2
    if Z1 == F2_ZERO:
        return j2
    if Z2 == F2_ZERO:
        return j1
    Z1Z1, Z2Z2 = f2_sqr(Z1), f2_sqr(Z2)
    U1, U2 = f2_mul(X1, Z2Z2), f2_mul(X2, Z1Z1)
    S1, S2 = f2_mul(f2_mul(Y1, Z2), Z2Z2), f2_mul(f2_mul(Y2, Z1), Z1Z1)
    if U1 == U2:
        if S1 == S2:
            return _g2_jac_double(j1)
        return (F2_ONE, F2_ONE, F2_ZERO)
    H = f2_sub(U2, U1)
    I = f2_sqr(f2_scale(H, 2))
    J = f2_mul(H, I)
    r = f2_scale(f2_sub(S2, S1), 2)
    V = f2_mul(U1, I)
    X3 = f2_sub(f2_sub(f2_sqr(r), J), f2_scale(V, 2))
    Y3 = f2_sub(f2_mul(r, f2_sub(V, X3)), f2_scale(f2_mul(S1, J), 2))
    Z3 = f2_mul(f2_sub(f2_sub(f2_sqr(f2_add(Z1, Z2)), Z1Z1), Z2Z2), H)
    return (X3, Y3, Z3)


def _g2_to_jac(p):
    return (F2_ONE, F2_ONE, F2_ZERO) if p is INF else (p[0], p[1], F2_ONE)


def _g2_from_jac(j):
    X, Y, Z = j
    if Z == F2_ZERO:
        return INF
    zi = f2_inv(Z)
    zi2 = f2_sqr(zi)
    return (f2_mul(X, zi2), f2_mul(Y, f2_mul(zi2, zi)))


def g2_add(p, q):
    return _g2_from_jac(_g2_jac_add(_g2_to_jac(p), _g2_to_jac(q)))


def g2_mul(p, k: int):
    """`mul` on G2 (/root/reference/src/Sonic/SRS.hs:35-36,40-41)."""
    k %= R
    if k == 0 or p is INF:
        return INF
    acc = (F2_ONE, F2_ONE, F2_ZERO)
    base = _g2_to_jac(p)
    for bit in bin(k)[2:]:
        acc = _g2_jac_double(acc)
        if bit == "1":
            acc = _g2_jac_add(acc, base)
    return _g2_from_jac(acc)


def g2_compress(p) -> bytes:
    """ZCash-style 96-byte compressed G2: x.c1 || x.c0 big-endian; flags as for G1; the sign bit
    is set when y is the lexicographically larger root (compare c1 first, then c0)."""
    if p is INF:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), (y0, y1) = p
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    b[0] |= 0x80
    big = (y1 > (Q - 1) // 2) if y1 != 0 else (y0 > (Q - 1) // 2)
    if big:
        b[0] |= 0x20
    return bytes(b)
