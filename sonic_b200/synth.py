"""Synthetic workloads: seeded field elements and circuits shaped like
the reference's generators (test/Test/Reference.hs:125-169).  Pure data generation -- no
field arithmetic beyond what is needed to make a circuit satisfiable (done with numpy-free
Python ints for small n, and a constant-weight construction for large n)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

R_MODULUS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
TRAPDOOR_SEED = 0x534F4E4943  # "SONIC"
_MASK = (1 << 64) - 1


def splitmix64(seed: int, count: int) -> np.ndarray:
    """`count` outputs of the SplitMix64 stream started at `seed` (vectorised)."""
    idx = np.arange(1, count + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _MASK) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def fr_ints(seed: int, count: int) -> List[int]:
    """Uniform-ish Fr values: 4 x 64 bits reduced mod r (exact, Python ints)."""
    w = splitmix64(seed, 4 * count).reshape(count, 4)
    out = []
    for a, b, c, d in w.tolist():
        out.append((a | (b << 64) | (c << 128) | (d << 192)) % R_MODULUS)
    return out


def fr_bytes_fast(seed: int, count: int) -> np.ndarray:
    """`count` canonical Fr encodings (count x 32 uint8) without big-int arithmetic: the top
    limb is masked to 62 bits, which keeps every value below r (r > 2^254)."""
    w = splitmix64(seed, 4 * count).reshape(count, 4).copy()
    w[:, 3] &= np.uint64((1 << 62) - 1)
    return w.view(np.uint8).reshape(count, 32)


def skewed_fr_bytes(seed: int, count: int) -> np.ndarray:
    """50% zeros, 25% +-1, the rest uniform."""
    b = fr_bytes_fast(seed, count).copy()
    sel = splitmix64(seed ^ 0xABCDEF, count) % np.uint64(8)
    zero = sel < 4
    one = (sel == 4)
    minus_one = (sel == 5)
    b[zero] = 0
    b[one] = 0
    b[one, 0] = 1
    m1 = np.frombuffer((R_MODULUS - 1).to_bytes(32, "little"), dtype=np.uint8)
    b[minus_one] = m1
    return b


def trapdoor() -> Tuple[int, int]:
    x, alpha = fr_ints(TRAPDOOR_SEED, 2)
    return x, alpha


def ints_to_bytes(xs) -> bytes:
    return b"".join((x % R_MODULUS).to_bytes(32, "little") for x in xs)


def synthetic_circuit(n: int, Q: int, seed: int):
    """Weights as the reference's generator makes them (test/Test/Reference.hs:141-149): per
    matrix one all-ones row among zero rows, at a seeded position; aL, aR uniform, aO = aL*aR,
    cs back-solved (:138).  Returns plain Python lists of ints."""
    sel = splitmix64(seed ^ 0x5EED, 3) % np.uint64(Q)
    rows = [int(v) for v in sel.tolist()]
    aL = fr_ints(seed * 3 + 1, n)
    aR = fr_ints(seed * 3 + 2, n)
    aO = [a * b % R_MODULUS for a, b in zip(aL, aR)]

    def mat(row):
        return [[1] * n if q == row else [0] * n for q in range(Q)]

    wL, wR, wO = mat(rows[0]), mat(rows[1]), mat(rows[2])
    sL, sR, sO = sum(aL) % R_MODULUS, sum(aR) % R_MODULUS, sum(aO) % R_MODULUS
    cs = [((sL if q == rows[0] else 0) + (sR if q == rows[1] else 0) + (sO if q == rows[2] else 0)) % R_MODULUS
          for q in range(Q)]
    return (wL, wR, wO, cs), (aL, aR, aO)


def synthetic_circuit_bytes(n: int, Q: int, seed: int):
    """Same circuit as `synthetic_circuit`, as byte buffers ready for the C ABI (no Q x n Python
    lists: n = 2^16 would be slow to build that way)."""
    sel = splitmix64(seed ^ 0x5EED, 3) % np.uint64(Q)
    rows = [int(v) for v in sel.tolist()]
    aL = fr_ints(seed * 3 + 1, n)
    aR = fr_ints(seed * 3 + 2, n)
    aO = [a * b % R_MODULUS for a, b in zip(aL, aR)]
    one = np.zeros(32, dtype=np.uint8)
    one[0] = 1

    def mat(row):
        m = np.zeros((Q, n, 32), dtype=np.uint8)
        m[row, :, :] = one
        return m

    sL, sR, sO = sum(aL) % R_MODULUS, sum(aR) % R_MODULUS, sum(aO) % R_MODULUS
    cs = [((sL if q == rows[0] else 0) + (sR if q == rows[1] else 0) + (sO if q == rows[2] else 0)) % R_MODULUS
          for q in range(Q)]
    return {
        "n": n, "Q": Q,
        "wL": mat(rows[0]), "wR": mat(rows[1]), "wO": mat(rows[2]),
        "cs": np.frombuffer(ints_to_bytes(cs), dtype=np.uint8).copy(),
        "aL": np.frombuffer(ints_to_bytes(aL), dtype=np.uint8).copy(),
        "aR": np.frombuffer(ints_to_bytes(aR), dtype=np.uint8).copy(),
        "aO": np.frombuffer(ints_to_bytes(aO), dtype=np.uint8).copy(),
        "ints": {"aL": aL, "aR": aR, "aO": aO, "cs": cs, "rows": rows},
    }
