"""CPU: pins the oracle.  The reference holds no golden vector, so the
oracle is anchored on public BLS12-381 vectors, on the trapdoor identities, on the
reference's own properties and circuits, and on the committed fixtures in tests/golden/."""
import json
import os
import random

import pytest

from oracle import bls12_381 as bls
from oracle import sonic as S
from tests.util import example1, example2, random_d, rnd_circuit

R, Q = bls.R, bls.Q
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_public_constants():
    z = -0xd201000000010000  # the BLS12-381 curve parameter
    assert R == z ** 4 - z ** 2 + 1
    assert Q == (z - 1) ** 2 * R // 3 + z
    assert (R - 1) % (1 << 32) == 0 and (R - 1) % (1 << 33) != 0
    assert pow(7, (R - 1) // 2, R) == R - 1  # 7 is a non-residue => generator of the 2-Sylow part
    assert bls.g1_is_on_curve(bls.G1_GEN)
    assert bls.g1_mul(bls.G1_GEN, R - 1) == bls.g1_neg(bls.G1_GEN)
    assert bls.g1_add(bls.g1_mul(bls.G1_GEN, R - 1), bls.G1_GEN) is bls.INF


def test_zcash_compressed_vectors():
    # identity, G and 2G of the ZCash/IETF serialisation (public test vectors)
    assert bls.g1_compress(bls.INF).hex() == "c0" + "00" * 47
    assert bls.g1_compress(bls.G1_GEN).hex() == (
        "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
        "6c55e83ff97a1aeffb3af00adb22c6bb")
    assert bls.g1_compress(bls.g1_mul(bls.G1_GEN, 2)).hex() == (
        "a572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62a"
        "e28f75bb8f1c7c42c39a8c5529bf0f4e")
    rng = random.Random(1)
    for _ in range(20):
        p = bls.g1_mul_gen(rng.randrange(R))
        assert bls.g1_decompress(bls.g1_compress(p)) == p
        assert bls.g1_from_raw(bls.g1_to_raw(p)) == p


def test_group_law_consistency():
    rng = random.Random(2)
    for _ in range(10):
        a, b = rng.randrange(R), rng.randrange(R)
        A, B = bls.g1_mul_gen(a), bls.g1_mul_gen(b)
        assert bls.g1_add(A, B) == bls.g1_mul_gen(a + b)
        assert bls.g1_mul(A, b) == bls.g1_mul_gen(a * b)
        assert bls.g1_mul(bls.G1_GEN, a) == A  # fixed-base table == double-and-add
    pts = [bls.g1_mul_gen(rng.randrange(R)) for _ in range(40)]
    scs = [rng.choice([0, 1, R - 1, rng.randrange(R)]) for _ in pts]
    assert S.msm_pippenger(pts, scs) == bls.g1_msm_naive(pts, scs)


def test_srs_index_conventions_and_trapdoor():
    """SRS.hs:33-39: which exponent sits at which index; g^alpha absent."""
    x, alpha, d = 11, 4, 6
    srs = S.srs_new(d, x, alpha)
    assert len(srs.gNegativeX) == d and len(srs.gPositiveX) == d + 1
    assert len(srs.gNegativeAlphaX) == d and len(srs.gPositiveAlphaX) == d
    xi = pow(x, -1, R)
    for i in range(1, d + 1):
        assert srs.gNegativeX[i - 1] == bls.g1_mul(bls.G1_GEN, pow(xi, i, R))
        assert srs.gPositiveAlphaX[i - 1] == bls.g1_mul(bls.G1_GEN, alpha * pow(x, i, R))
        assert srs.gNegativeAlphaX[i - 1] == bls.g1_mul(bls.G1_GEN, alpha * pow(xi, i, R))
    assert srs.gPositiveX[0] == bls.G1_GEN
    assert bls.g1_mul(bls.G1_GEN, alpha) not in srs.gPositiveAlphaX[:1]


def test_commit_open_trapdoor_identities():
    """commitPoly srs max f == g^(alpha x^(d-max) f(x));  W == g^((f(x)-f(z))/(x-z))."""
    rng = random.Random(3)
    x, alpha, d = rng.randrange(1, R), rng.randrange(1, R), 14
    srs = S.srs_new(d, x, alpha)
    for maxm in (5, 14):
        f = {e: rng.randrange(R) for e in range(-maxm, maxm + 1) if e != -(d - maxm)}
        F = S.commitPoly(srs, maxm, f)
        fx = S.l_eval(f, x)
        assert F == bls.g1_mul_gen(alpha * pow(x, d - maxm, R) * fx)
        z = rng.randrange(1, R)
        v, W = S.openPoly(srs, z, f)
        assert v == S.l_eval(f, z)
        assert W == bls.g1_mul_gen((fx - v) * pow(x - z, -1, R))
        assert S.pcV_trapdoor(srs, maxm, F, z, (v, W))
        assert not S.pcV_trapdoor(srs, maxm, F, z, ((v + 1) % R, W))


def test_reference_constraint_properties():
    """test/Test/Constraints.hs: r(X,Y) = r(XY,1) (:30-34); zero constant terms (:37-83)."""
    rng = random.Random(4)
    for _ in range(5):
        circuit, assignment = rnd_circuit(rng, n=rng.randint(1, 8))
        n = len(assignment.aL)
        rXY = S.rPoly(assignment)
        sXY = S.sPoly(circuit.weights)
        x, y = rng.randrange(1, R), rng.randrange(1, R)
        assert S.l_eval(S.evalY(y, rXY), x) == S.l_eval(S.evalY(1, rXY), x * y % R)
        assert 0 not in rXY and 0 not in sXY and 0 not in S.bv_add(rXY, sXY)
        tXY = S.tPoly(rXY, sXY, S.kPoly(circuit.cs, n))
        assert 0 not in tXY  # satisfied circuit <=> X^0 coefficient of t vanishes (:66-83)
        # a^T w_L + b^T w_R + c^T w_O = k (:19-27)
        w = circuit.weights
        for q, k in enumerate(circuit.cs):
            dot = lambda v, row: sum(a * b for a, b in zip(v, row))
            assert (dot(assignment.aL, w.wL[q]) + dot(assignment.aR, w.wR[q]) + dot(assignment.aO, w.wO[q])) % R == k


@pytest.mark.parametrize("which", ["example1", "example2"])
def test_prove_verify_reference_circuits(which):
    """examples/Main.hs / test_sonic: verify (prove ...) == True; literal == dense layer."""
    rng = random.Random(5)
    circuit, assignment = example1() if which == "example1" else example2(12)
    n, Qn = len(assignment.aL), len(circuit.weights.wL)
    d = 12 if n == 1 else 16
    srs = S.srs_new(d, rng.randrange(1, R), rng.randrange(1, R))
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Qn))]
    proof, (y, z, yzs) = S.prove(srs, assignment, circuit, rnd)
    assert S.verify_trapdoor(srs, circuit, proof, y, z, yzs)
    dense, _ = S.prove_dense(srs, assignment, circuit, rnd)
    assert S.encode_proof(dense) == S.encode_proof(proof)
    assert len(S.encode_proof(proof)) == S.proof_size(Qn)
    assert S.decode_proof(S.encode_proof(proof), Qn) == proof
    # a forged evaluation must fail
    proof.prA = (proof.prA + 1) % R
    assert not S.verify_trapdoor(srs, circuit, proof, y, z, yzs)


def test_prove_random_circuit_dense_equals_literal():
    rng = random.Random(6)
    circuit, assignment = rnd_circuit(rng, n=4, m=3)
    d = random_d(rng, 4)
    srs = S.srs_new(d, rng.randrange(1, R), rng.randrange(1, R))
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(3))]
    a, (y, z, yzs) = S.prove(srs, assignment, circuit, rnd)
    b, _ = S.prove_dense(srs, assignment, circuit, rnd)
    assert S.encode_proof(a) == S.encode_proof(b)
    assert S.verify_trapdoor(srs, circuit, b, y, z, yzs)


def test_panics_match_reference_text():
    circuit, assignment = example2(12)
    rnd = list(range(1, S.rnd_count(5) + 1))
    with pytest.raises(S.SonicPanic, match="Parameter d is not large enough: 13 should be greater than 14"):
        S.prove(S.srs_new(13, 3, 5), assignment, circuit, rnd)
    # README's d >= 3n+9 = 15 is below what the code needs (max(7n, 4n+8) = 16): SURVEY.md section 8d config 1
    with pytest.raises(S.SonicPanic, match="commitPoly: gNegativeAlphaX is not long enough: 15 >= 15"):
        S.prove(S.srs_new(15, 3, 5), assignment, circuit, rnd)
    srs = S.srs_new(8, 3, 5)
    with pytest.raises(S.SonicPanic, match="gNegativeAlphaX is not long enough: -1 >= 8"):
        S.commitPoly(srs, 8, {0: 1})
    with pytest.raises(S.SonicPanic, match="openPoly: gPositiveX is not long enough: 9 >= 9"):
        S.openPoly(srs, 2, {10: 1})


def test_golden_fixtures():
    """Fixtures generated by tools/gen_golden.py from this oracle; they freeze its outputs so
    that a later change to the oracle cannot silently move the parity target."""
    with open(os.path.join(GOLDEN, "prove_small.json")) as fh:
        cases = json.load(fh)
    assert cases
    for c in cases:
        circuit = S.ArithCircuit(S.GateWeights(c["wL"], c["wR"], c["wO"]), c["cs"])
        assignment = S.Assignment(c["aL"], c["aR"], c["aO"])
        srs = S.srs_new(c["d"], c["x"], c["alpha"])
        proof, _ = S.prove_dense(srs, assignment, circuit, c["rnd"])
        assert S.encode_proof(proof).hex() == c["proof_hex"], c["name"]
    with open(os.path.join(GOLDEN, "commit_open.json")) as fh:
        cases = json.load(fh)
    for c in cases:
        srs = S.srs_new(c["d"], c["x"], c["alpha"])
        f = {int(e): v for e, v in c["f"].items()}
        assert bls.g1_compress(S.commitPoly(srs, c["max"], f)).hex() == c["commit_hex"]
        v, w = S.openPoly(srs, c["z"], f)
        assert (v, bls.g1_compress(w).hex()) == (c["value"], c["open_hex"])
