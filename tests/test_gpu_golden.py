"""GPU: the committed fixtures under tests/golden/ replayed through the C ABI, BASELINE config 1's d sweep
against the C restatement, and the round-2 entry points (general hscProve, batches, restricted window
tables, uniform verdicts of sharded proofs)."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

from oracle import bls12_381 as bls
from oracle import cref
from oracle import sonic as S
from sonic_b200 import synth
from tests.util import example1, example2, rnd_circuit, to_gpu_types

pytestmark = pytest.mark.gpu
R = bls.R
C = bls.g1_compress
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_proofs_replayed_through_the_abi(gpu):
    """tests/golden/prove_small.json (tools/gen_golden.py: the reference's two fixed circuits at random and at the
    bench trapdoor x = 1, alpha = 4, d = 25n of bench/Main.hs:18-27, and two generator-shaped circuits): the CUDA
    proof equals the frozen bytes, dense and CSR ingestion alike."""
    cases = json.load(open(os.path.join(GOLDEN, "prove_small.json")))
    assert len(cases) >= 6
    for c in cases:
        srs = gpu.SRS.new(c["d"], c["x"], c["alpha"])
        for sparse in (False, True):
            circ = gpu.ArithCircuit(gpu.GateWeights(c["wL"], c["wR"], c["wO"]), c["cs"], sparse=sparse)
            got = gpu.prove_bytes(srs, gpu.Assignment(c["aL"], c["aR"], c["aO"]), circ, c["rnd"])
            assert got.hex() == c["proof_hex"], (c["name"], sparse)
        srs.free()


def test_golden_commit_open_replayed_through_the_abi(gpu):
    """tests/golden/commit_open.json: commitPoly / openPoly on sparse Laurent polynomials with +-1 coefficients."""
    cases = json.load(open(os.path.join(GOLDEN, "commit_open.json")))
    assert len(cases) >= 3
    for c in cases:
        srs = gpu.SRS.new(c["d"], c["x"], c["alpha"])
        f = {int(e): v for e, v in c["f"].items()}
        assert gpu.commitPoly(srs, c["max"], f).hex() == c["commit_hex"]
        v, w = gpu.openPoly(srs, c["z"], f)
        assert v == c["value"] and w.hex() == c["open_hex"]
        srs.free()


def _cref_prove(circuit, assignment, d, x, alpha, rnd):
    w = circuit.weights
    flat = lambda m: np.frombuffer(synth.ints_to_bytes([v for row in m for v in row]), dtype=np.uint8).copy()
    vec = lambda v: np.frombuffer(synth.ints_to_bytes(v), dtype=np.uint8).copy()
    table = cref.srs_new(d, x, alpha, threads=4)
    n, Q = len(assignment.aL), len(w.wL)
    return cref.prove(table, d, n, Q, flat(w.wL), flat(w.wR), flat(w.wO), vec(circuit.cs), vec(assignment.aL), vec(assignment.aR),
                      vec(assignment.aO), vec(rnd), threads=4)


def test_config1_d_sweep_16_to_200(gpu):
    """BASELINE config 1: examples/Main.hs:38-63 (= arithCircuitExample2, n = 2, Q = 5), SRS.new + prove for EVERY d in
    [16, 200] (randomD's range for n = 2, test/Test/Reference.hs:101-104) -- random trapdoor on even d, the bench
    trapdoor x = 1, alpha = 4 (bench/Main.hs:18-27: every base coincides, P + P and P - P occur) on odd d -- byte for byte
    against the C restatement; d = 15 (README's 3n+9) reproduces the reference panic instead of a proof."""
    rng = random.Random(2024)
    circuit, assignment = example2(12)
    gc, ga = to_gpu_types(gpu, circuit, assignment)
    for d in range(16, 201):
        x, alpha = (1, 4) if d % 2 else (rng.randrange(1, R), rng.randrange(1, R))
        rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(5))]
        srs = gpu.SRS.new(d, x, alpha)
        got = gpu.prove_bytes(srs, ga, gc, rnd)
        assert got == _cref_prove(circuit, assignment, d, x, alpha, rnd), d
        srs.free()
    # the same two circuits at the bench shape d = 25 n with x = 1, alpha = 4, against the Python oracle and verified
    for circuit_, assignment_ in (example1(), example2(12)):
        n, Q = len(assignment_.aL), len(circuit_.weights.wL)
        d = 25 * n
        rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
        o = S.srs_new(d, 1, 4)
        want, (y, z, yzs) = S.prove(o, assignment_, circuit_, rnd)
        gc_, ga_ = to_gpu_types(gpu, circuit_, assignment_)
        got = gpu.prove_bytes(gpu.SRS.new(d, 1, 4), ga_, gc_, rnd)
        assert got == S.encode_proof(want)
        assert S.verify_trapdoor(o, circuit_, S.decode_proof(got, Q), y, z, yzs)
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(5))]
    with pytest.raises(S.SonicPanic) as oe:
        S.prove(S.srs_new(15, 3, 5), assignment, circuit, rnd)
    with pytest.raises(gpu.SonicError) as ge:
        gpu.prove_bytes(gpu.SRS.new(15, 3, 5), ga, gc, rnd)
    assert ge.value.text == str(oe.value) and "is not long enough" in ge.value.text


def test_sharded_proof_fails_uniformly_on_an_unsatisfied_assignment(gpu):
    """An assignment that does not satisfy the circuit leaves a non-zero X^0 term in t(X,y): commitPoly indexes g^alpha
    and the reference panics (CommitmentScheme.hs:70-73).  Only the ranks that build t see the coefficient; the
    verdict travels in the exchange record, so EVERY rank's shard call succeeds and EVERY fold reports the panic."""
    rng = random.Random(31)
    circuit, assignment = rnd_circuit(rng, 9, 4)
    n, Q = 9, 4
    d = 7 * n + 2
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
    g = gpu.SRS.new(d, x, alpha)
    gc, _ = to_gpu_types(gpu, circuit, assignment)
    bad = gpu.Assignment(list(assignment.aL), list(assignment.aR), [(assignment.aO[0] + 1) % R] + list(assignment.aO[1:]))
    with pytest.raises(gpu.SonicError) as single:
        gpu.prove_bytes(g, bad, gc, rnd)
    assert single.value.kind == "SRS_TOO_SHORT" and single.value.text == "commitPoly: gNegativeAlphaX is not long enough: -1 >= %d" % d
    for world in (2, 3, 5, 8):
        blobs = [gpu.prove_shard(g, bad, gc, rnd, r, world) for r in range(world)]   # no rank raises here
        with pytest.raises(gpu.SonicError) as e:
            gpu.prove_combine(Q, blobs)
        assert (e.value.kind, e.value.text) == (single.value.kind, single.value.text), world
    # and a satisfied one folds to the single-GPU proof for the same rank counts
    _, ga = to_gpu_types(gpu, circuit, assignment)
    want = gpu.prove_bytes(g, ga, gc, rnd)
    o = S.srs_new(d, x, alpha)
    assert want == S.encode_proof(S.prove_dense(o, assignment, circuit, rnd)[0])
    for world in (2, 3, 5, 8):
        assert gpu.prove_combine(Q, [gpu.prove_shard(g, ga, gc, rnd, r, world) for r in range(world)]) == want, world


def test_hsc_prove_on_a_general_bivariate_polynomial(gpu):
    """test/Test/Signature.hs:20-36 as written there: `hscProve srs (sPoly weights) yzs` -- the TERMS of s(X,Y), not
    the weights, cross the ABI (sonic_hsc_prove_terms); then polynomials no circuit produces (repeated monomials,
    zero coefficients, only non-negative exponents, the zero polynomial)."""
    rng = random.Random(41)

    def same(got, want, u, v):
        assert [(a, (b, c)) for a, (b, c) in got.hscS] == [(C(a), (b, C(c))) for a, (b, c) in want.hscS]
        assert got.hscW == [(a, C(b), C(c)) for a, b, c in want.hscW]
        assert (got.hscQv, got.hscC, got.hscU, got.hscV) == (C(want.hscQv), C(want.hscC), u, v)

    for trial in range(3):
        circuit, assignment = rnd_circuit(rng, n=rng.randint(2, 9))
        n, Q = len(assignment.aL), len(circuit.weights.wL)
        d = max(7 * n, rng.randint(7 * n, 20 * n))
        x, alpha = rng.randrange(1, R), rng.randrange(1, R)
        g, o = gpu.SRS.new(d, x, alpha), S.srs_new(d, x, alpha)
        m = Q + trial
        yzs = [(rng.randrange(1, R), rng.randrange(1, R)) for _ in range(m)]
        u, v = rng.randrange(1, R), rng.randrange(1, R)
        sXY = S.sPoly(circuit.weights)
        want = S.hscProve(o, sXY, yzs, u, v)
        same(gpu.hscProveBiV(g, sXY, yzs, u, v), want, u, v)
        # the circuit-handle entry gives the same bytes
        gc, _ = to_gpu_types(gpu, circuit, assignment)
        assert gpu.hscProveBiV(g, sXY, yzs, u, v) == gpu.hscProve(g, gc, yzs, u, v)
        assert S.hscVerify_trapdoor(o, sXY, yzs, want)
    # arbitrary sparse polynomials
    d = 40
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    g, o = gpu.SRS.new(d, x, alpha), S.srs_new(d, x, alpha)
    for kind in ("mixed", "nonnegative", "zero"):
        sXY = {}
        if kind != "zero":
            for _ in range(25):
                ex = rng.randint(0 if kind == "nonnegative" else -12, 14)
                ey = rng.randint(0 if kind == "nonnegative" else -9, 11)
                if ex == 0:
                    continue   # commitPoly srs d: a term at X^0 indexes g^alpha (tested below)
                sXY.setdefault(ex, {})[ey] = rng.choice([1, R - 1, rng.randrange(R)])
        yzs = [(rng.randrange(1, R), rng.randrange(1, R)) for _ in range(3)]
        u, v = rng.randrange(1, R), rng.randrange(1, R)
        sn = S.bv_norm(sXY) if hasattr(S, "bv_norm") else sXY
        # s(u, Y) is committed with max = d too: its Y^0 coefficient must vanish for the reference not to panic
        if any(0 in inner for inner in sn.values()):
            for inner in sn.values():
                inner.pop(0, None)
            sn = {e: c for e, c in sn.items() if c}
        want = S.hscProve(o, sn, yzs, u, v)
        same(gpu.hscProveBiV(g, sn, yzs, u, v), want, u, v)
    # a constant term in X: the reference's `index` panic, same text
    sXY = {0: {3: 5}, 2: {1: 7}}
    yzs = [(3, 4)]
    with pytest.raises(S.SonicPanic) as oe:
        S.hscProve(o, sXY, yzs, 5, 6)
    with pytest.raises(gpu.SonicError) as ge:
        gpu.hscProveBiV(g, sXY, yzs, 5, 6)
    assert ge.value.text == str(oe.value)


def test_prove_batch_equals_single_proofs(gpu):
    rng = random.Random(51)
    circuit, assignment = rnd_circuit(rng, 7, 3)
    n, Q = 7, 3
    d = 60
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    g = gpu.SRS.new(d, x, alpha)
    gc, ga = to_gpu_types(gpu, circuit, assignment)
    rnds = [[rng.randrange(1, R) for _ in range(S.rnd_count(Q))] for _ in range(5)]
    singles = [gpu.prove_bytes(g, ga, gc, r) for r in rnds]
    assert gpu.prove_batch(g, [ga] * 5, gc, rnds) == singles
    assert gpu.prove_batch(g, [], gc, []) == []
    o = S.srs_new(d, x, alpha)
    assert singles[0] == S.encode_proof(S.prove_dense(o, assignment, circuit, rnds[0])[0])


def test_window_tables_restricted_to_the_circuit_size(gpu):
    """When the full-range window tables exceed the memory budget, the first proof of a circuit size builds tables for the
    17n+23 exponents that size reads (by doubling; no trapdoor).  Same proof bytes as with full tables and as without any."""
    from sonic_b200 import capi

    n, Q = 1 << 9, 4
    d = 40 * n   # much longer than 7n: the restricted ranges are a small part of it
    x, alpha = synth.trapdoor()
    c = synth.synthetic_circuit_bytes(n, Q, seed=9)
    rnd = [v or 1 for v in synth.fr_ints(91, 2 * Q + 8)]
    L = capi.lib()

    def prove_with(budget_mb, precompute):
        gpu.set_option("precompute_budget_mb", budget_mb)
        gpu.set_option("precompute", precompute)
        try:
            srs = gpu.SRS.new(d, x, alpha)
            ch = ctypes.c_void_p()
            capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
            out = ctypes.create_string_buffer(int(L.sonic_proof_size(Q)))
            w = ctypes.c_uint64(0)
            rb = np.frombuffer(synth.ints_to_bytes(rnd), dtype=np.uint8).copy()
            for _ in range(2):   # the second call reuses the cached tables
                capi.check(L.sonic_prove(srs._h, ch, c["aL"].ctypes.data, c["aR"].ctypes.data, c["aO"].ctypes.data, rb.ctypes.data, out, len(out), ctypes.byref(w)))
            pre = gpu.last_timing_ms("msm.precomputed")
            # standalone calls on the same handle: inside and outside the restricted ranges
            f = {e: (e * 7 + 3) % R for e in range(-n, n + 1) if e}
            cm = gpu.commitPoly(srs, d, f)
            far = gpu.msm(srs, 0, 10 * n, [5, 6, 7])
            L.sonic_circuit_free(ch)
            srs.free()
            return out.raw, pre, cm, far
        finally:
            gpu.set_option("precompute_budget_mb", 8192)
            gpu.set_option("precompute", -1)

    full = prove_with(8192, -1)
    restricted = prove_with(64, -1)      # 4d+2 points x 16 levels do not fit 64 MB; 17n points do
    none = prove_with(8192, 0)
    assert full[1] == 1.0 and restricted[1] == 1.0 and none[1] == 0.0
    assert full[0] == restricted[0] == none[0]
    assert full[2:] == restricted[2:] == none[2:]
    table = cref.srs_new(d, x, alpha, threads=8)
    want = cref.prove(table, d, n, Q, c["wL"], c["wR"], c["wO"], c["cs"], c["aL"], c["aR"], c["aO"],
                      np.frombuffer(synth.ints_to_bytes(rnd), dtype=np.uint8).copy(), threads=8)
    assert full[0] == want


def test_far_exponents_and_damaged_srs_files(gpu, tmp_path):
    """a short polynomial at a far exponent must give the reference's `index` panic from host-side bounds
    (not an out-of-memory error, not a truncated window); an SRS file whose header breaks the invariants SRS.new
    enforces, or whose size disagrees with its header, is refused."""
    import struct

    from sonic_b200 import capi

    rng = random.Random(61)
    d = 20
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    g, o = gpu.SRS.new(d, x, alpha), S.srs_new(d, x, alpha)
    for e in (10 ** 9, -10 ** 9, 2 ** 40, -2 ** 40, 2 ** 62):
        f = {e: 5}
        with pytest.raises(gpu.SonicError) as ge:
            gpu.commitPoly(g, d, f)
        assert ge.value.kind == "SRS_TOO_SHORT"
        if abs(e) <= 10 ** 9:
            with pytest.raises(S.SonicPanic) as oe:
                S.commitPoly(o, d, f)
            assert ge.value.text == str(oe.value)
        with pytest.raises(gpu.SonicError) as ge:
            gpu.openPoly(g, 7, f)
        assert ge.value.kind == "SRS_TOO_SHORT" and "is not long enough" in ge.value.text
    # zero coefficients far away are not terms: the polynomial below is 3 X^2
    f = {2: 3, 10 ** 12: 0, -10 ** 12: 0}
    assert gpu.commitPoly(g, d, f) == C(S.commitPoly(o, d, {2: 3}))
    v, w = gpu.openPoly(g, 9, f)
    vo, wo = S.openPoly(o, 9, {2: 3})
    assert (v, w) == (vo, C(wo))
    # z = 0 with a negative power: recip 0, whatever zeros pad the window
    with pytest.raises(gpu.SonicError) as ge:
        gpu.openPoly(g, 0, {-1: 1, 3: 2})
    assert ge.value.kind == "DIV_BY_ZERO"
    # SRS files
    path = str(tmp_path / "srs.bin")
    g.save(path)
    blob = open(path, "rb").read()
    assert gpu.SRS.load(path).gPositiveX == g.gPositiveX
    magic, version, pre_c, dd, levels, ppl = struct.unpack_from("<8sIIQQQ", blob)
    assert magic == b"SONICSRS" and dd == d and ppl == 2 * (2 * d + 1)

    def refused(data):
        bad = str(tmp_path / "bad.bin")
        open(bad, "wb").write(data)
        with pytest.raises(gpu.SonicError) as e:
            gpu.SRS.load(bad)
        assert e.value.kind == "INVALID_ARG"

    refused(blob[:-1])                                                              # truncated
    refused(blob + b"\0")                                                           # longer than its header says
    refused(blob[:20])                                                              # shorter than the header
    refused(struct.pack("<8sIIQQQ", magic, version, 255, dd, 1, ppl) + blob[40:])   # window bits out of range
    refused(struct.pack("<8sIIQQQ", magic, version, pre_c, dd, levels + 1, ppl) + blob[40:])
    refused(struct.pack("<8sIIQQQ", b"NOTANSRS", version, pre_c, dd, levels, ppl) + blob[40:])
