"""CPU: the C restatement (oracle/csrc/sonic_ref.c) against the Python big-int oracle and the
committed fixtures.  Both follow the reference's algorithm; they must agree byte for byte."""
import json
import os
import random

import pytest

from oracle import bls12_381 as bls
from oracle import cref
from oracle import sonic as S
from tests.util import example2, rnd_circuit

R = bls.R
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
fb = lambda xs: b"".join((x % R).to_bytes(32, "little") for x in xs)


def test_c_srs_and_msm_match_python_oracle():
    rng = random.Random(21)
    d, x, alpha = 9, rng.randrange(1, R), rng.randrange(1, R)
    raw = cref.srs_new(d, x, alpha, threads=2)
    o = S.srs_new(d, x, alpha)
    stride = 2 * d + 1
    pt = lambda fam, k: bls.g1_from_raw(raw[96 * (fam * stride + k + d):96 * (fam * stride + k + d) + 96])
    assert [pt(0, k) for k in range(0, d + 1)] == o.gPositiveX
    assert [pt(0, -k) for k in range(1, d + 1)] == o.gNegativeX
    assert [pt(1, k) for k in range(1, d + 1)] == o.gPositiveAlphaX
    assert [pt(1, -k) for k in range(1, d + 1)] == o.gNegativeAlphaX
    assert pt(1, 0) is bls.INF
    sc = [rng.choice([0, 1, R - 1, rng.randrange(R)]) for _ in range(stride)]
    want = bls.g1_msm_naive([pt(0, k) for k in range(-d, d + 1)], sc)
    for threads in (1, 3):
        assert cref.msm_naive(raw[:96 * stride], fb(sc), stride, threads) == bls.g1_compress(want)
    for c in (4, 7, 13):
        assert cref.msm_pippenger(raw[:96 * stride], fb(sc), stride, c, threads=3) == bls.g1_compress(want)
    with pytest.raises(ZeroDivisionError):
        cref.srs_new(4, 0, 1)


def test_c_commit_open_match_golden():
    with open(os.path.join(GOLDEN, "commit_open.json")) as fh:
        cases = json.load(fh)
    for c in cases:
        raw = cref.srs_new(c["d"], c["x"], c["alpha"], threads=4)
        f = {int(e): v for e, v in c["f"].items()}
        lo, hi = min(min(f), 0), max(max(f), 0)
        dense = fb([f.get(e, 0) for e in range(lo, hi + 1)])
        assert cref.commit(raw, c["d"], c["max"], lo, dense, threads=2).hex() == c["commit_hex"]
        v, w = cref.open_(raw, c["d"], c["z"], lo, dense, threads=2)
        assert (v, w.hex()) == (c["value"], c["open_hex"])
    raw = cref.srs_new(8, 3, 5)
    with pytest.raises(cref.RefPanic) as e:
        cref.commit(raw, 8, 8, 0, fb([1]))
    assert (e.value.code, e.value.exponent) == (2, 0)


def test_c_prove_matches_golden_and_python():
    with open(os.path.join(GOLDEN, "prove_small.json")) as fh:
        cases = json.load(fh)
    for c in cases:
        n, Q = len(c["aL"]), len(c["wL"])
        raw = cref.srs_new(c["d"], c["x"], c["alpha"], threads=4)
        flat = lambda m: fb([v for row in m for v in row])
        got = cref.prove(raw, c["d"], n, Q, flat(c["wL"]), flat(c["wR"]), flat(c["wO"]), fb(c["cs"]),
                         fb(c["aL"]), fb(c["aR"]), fb(c["aO"]), fb(c["rnd"]), threads=4)
        assert got.hex() == c["proof_hex"], c["name"]
    circuit, assignment = example2(12)
    raw = cref.srs_new(13, 3, 5)
    flat = lambda m: fb([v for row in m for v in row])
    w = circuit.weights
    with pytest.raises(cref.RefPanic) as e:
        cref.prove(raw, 13, 2, 5, flat(w.wL), flat(w.wR), flat(w.wO), fb(circuit.cs), fb(assignment.aL),
                   fb(assignment.aR), fb(assignment.aO), fb(list(range(1, 19))))
    assert e.value.code == 3
