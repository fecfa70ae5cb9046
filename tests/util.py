"""Shared helpers for the tests: the reference's fixed circuits and generators
(test/Test/Reference.hs), restated as data for the oracle and for the CUDA path."""
from __future__ import annotations

import random
from typing import Tuple

from oracle import bls12_381 as bls
from oracle import sonic as S

R = bls.R


def example1():
    """arithCircuitExample1 (test/Test/Reference.hs:38-50): n = 1, Q = 2."""
    wL, wR, wO = [[1], [0]], [[0], [1]], [[0], [0]]
    cs = [7 + 3, 2 + 10]
    aL, aR = [10], [12]
    aO = [a * b % R for a, b in zip(aL, aR)]
    return S.ArithCircuit(S.GateWeights(wL, wR, wO), cs), S.Assignment(aL, aR, aO)


def example2(z: int = 12):
    """arithCircuitExample2 (test/Test/Reference.hs:65-90) = examples/Main.hs:38-63: n = 2, Q = 5."""
    f = lambda m: [[v % R for v in row] for row in m]
    wL = f([[0, 0], [1, 0], [0, 1], [0, 0], [0, 0]])
    wR = f([[0, 0], [0, 0], [0, 0], [1, 0], [0, 1]])
    wO = f([[1, -1], [0, 0], [0, 0], [0, 0], [0, 0]])
    cs = [c % R for c in [0, 4 - z, 9 - z, 9 - z, 4 - z]]
    aL = [(4 - z) % R, (9 - z) % R]
    aR = [(9 - z) % R, (4 - z) % R]
    aO = [a * b % R for a, b in zip(aL, aR)]
    return S.ArithCircuit(S.GateWeights(wL, wR, wO), cs), S.Assignment(aL, aR, aO)


def rnd_circuit(rng: random.Random, n: int | None = None, m: int | None = None):
    """rndCircuit (test/Test/Reference.hs:125-169): n in [1,20], Q in [1,n]; each weight matrix is
    one all-ones row inserted among Q-1 zero rows; cs back-solved from the assignment."""
    n = n or rng.randint(1, 20)
    m = m or rng.randint(1, n)
    aL = [rng.randrange(R) for _ in range(n)]
    aR = [rng.randrange(R) for _ in range(n)]
    aO = [a * b % R for a, b in zip(aL, aR)]

    def gen():
        i = min(rng.randint(0, m), m - 1)
        rows = [[0] * n for _ in range(m - 1)]
        rows.insert(i, [1] * n)
        return rows

    wL, wR, wO = gen(), gen(), gen()
    dot = lambda v, mat: [sum(a * b for a, b in zip(v, row)) % R for row in mat]
    cs = [(a + b + c) % R for a, b, c in zip(dot(aL, wL), dot(aR, wR), dot(aO, wO))]
    return S.ArithCircuit(S.GateWeights(wL, wR, wO), cs), S.Assignment(aL, aR, aO)


def random_d(rng: random.Random, n: int) -> int:
    """randomD (test/Test/Reference.hs:101-104)."""
    if n == 1:
        return rng.randint(12, 100)
    if n == 2:
        return rng.randint(16, 200)
    return rng.randint(7 * n, 100 * n)


def to_gpu_types(sb, circuit: S.ArithCircuit, assignment: S.Assignment):
    w = circuit.weights
    return (sb.ArithCircuit(sb.GateWeights(w.wL, w.wR, w.wO), circuit.cs),
            sb.Assignment(assignment.aL, assignment.aR, assignment.aO))


def limbs(v: int, n: int):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def from_limbs(a) -> int:
    return sum(int(x) << (32 * i) for i, x in enumerate(a))
