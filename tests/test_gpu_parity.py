"""GPU parity: SRS.new, commitPoly, openPoly, msm, prove and hscProve through the C ABI
against the CPU oracle on the same seeded inputs.  Byte-exact."""
import random

import pytest

from oracle import bls12_381 as bls
from oracle import sonic as S
from tests.util import example1, example2, random_d, rnd_circuit, to_gpu_types

pytestmark = pytest.mark.gpu
R = bls.R
C = bls.g1_compress


def _srs_pair(sb, d, x, alpha):
    return sb.SRS.new(d, x, alpha), S.srs_new(d, x, alpha)


def test_srs_new_matches_reference_vectors(gpu):
    rng = random.Random(11)
    for d in (1, 2, 7, 33, 130):
        x, alpha = rng.randrange(1, R), rng.randrange(R)
        g, o = _srs_pair(gpu, d, x, alpha)
        assert g.gNegativeX == [C(p) for p in o.gNegativeX]
        assert g.gPositiveX == [C(p) for p in o.gPositiveX]
        assert g.gNegativeAlphaX == [C(p) for p in o.gNegativeAlphaX]
        assert g.gPositiveAlphaX == [C(p) for p in o.gPositiveAlphaX]


def test_srs_degenerate_trapdoors(gpu):
    # bench/Main.hs:23 uses x = 1 (every base equals g or g^alpha); alpha = 0 makes a family infinity
    for x, alpha in ((1, 4), (R - 1, 1), (5, 0)):
        g, o = _srs_pair(gpu, 9, x, alpha)
        assert g.gPositiveX == [C(p) for p in o.gPositiveX]
        assert g.gNegativeAlphaX == [C(p) for p in o.gNegativeAlphaX]
    with pytest.raises(gpu.SonicError) as e:
        gpu.SRS.new(4, 0, 3)
    assert e.value.kind == "DIV_BY_ZERO"
    with pytest.raises(gpu.SonicError) as e:
        g.g1(1, 0)  # g^alpha is not part of the SRS (SRS.hs:38)
    assert e.value.kind == "SRS_TOO_SHORT"


def _rand_laurent(rng, lo, hi, density=0.8):
    f = {e: rng.randrange(R) for e in range(lo, hi + 1) if rng.random() < density}
    return f


def test_commit_open_small_vs_oracle(gpu):
    rng = random.Random(12)
    for trial in range(6):
        d = rng.randint(20, 60)
        x, alpha = (1, 4) if trial == 0 else (rng.randrange(1, R), rng.randrange(1, R))
        g, o = _srs_pair(gpu, d, x, alpha)
        maxm = rng.randint(5, d)
        f = _rand_laurent(rng, -maxm, maxm)  # shifted by d - max: stays inside [-d, d]
        f.pop(-(d - maxm), None)  # shifted exponent 0 is the alpha hole
        assert gpu.commitPoly(g, maxm, f) == C(S.commitPoly(o, maxm, f))
        z = rng.randrange(1, R)
        v, w = gpu.openPoly(g, z, f)
        ov, ow = S.openPoly(o, z, f)
        assert (v, w) == (ov, C(ow))
        assert S.pcV_trapdoor(o, maxm, bls.g1_decompress(gpu.commitPoly(g, maxm, f)), z, (v, bls.g1_decompress(w)))


def test_commit_open_edge_cases(gpu):
    rng = random.Random(13)
    d = 24
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    g, o = _srs_pair(gpu, d, x, alpha)
    # empty polynomial, constants, single terms, small coefficients (+-1, as in Constraints.hs:45)
    cases = [{}, {3: 0}, {0: 5}, {1: 1}, {-1: R - 1}, {-d: 7, d: 9}, {k: 1 for k in range(-d, d + 1) if k},
             {k: R - 1 for k in range(-5, 6)}]
    for f in cases:
        z = rng.randrange(1, R)
        v, w = gpu.openPoly(g, z, f)
        ov, ow = S.openPoly(o, z, f)
        assert (v, w) == (ov, C(ow)), f
    for f in cases:
        if any(e == 0 and c % R for e, c in f.items()):
            continue
        assert gpu.commitPoly(g, d, f) == C(S.commitPoly(o, d, f)), f
    # panics of `index` (CommitmentScheme.hs:70-73), text included
    for maxm, f in ((d, {0: 1}), (d, {d + 1: 1}), (d, {-d - 1: 2, 3: 1}), (d - 3, {d: 1})):
        with pytest.raises(S.SonicPanic) as oe:
            S.commitPoly(o, maxm, f)
        with pytest.raises(gpu.SonicError) as ge:
            gpu.commitPoly(g, maxm, f)
        assert ge.value.kind == "SRS_TOO_SHORT" and ge.value.text == str(oe.value)
    for f in ({d + 2: 1}, {-d - 2: 1, 0: 3}):
        with pytest.raises(S.SonicPanic) as oe:
            S.openPoly(o, 5, f)
        with pytest.raises(gpu.SonicError) as ge:
            gpu.openPoly(g, 5, f)
        assert ge.value.text == str(oe.value)
    # z = 0: fine without negative powers, `recip 0` with them
    assert gpu.openPoly(g, 0, {0: 4, 2: 9}) == (4, C(S.openPoly(o, 0, {0: 4, 2: 9})[1]))
    with pytest.raises(gpu.SonicError) as ge:
        gpu.openPoly(g, 0, {-1: 4, 2: 9})
    assert ge.value.kind == "DIV_BY_ZERO"


def test_msm_sizes_and_windows_vs_trapdoor(gpu):
    """MSM of N points against the trapdoor identity  sum_k s_k g^{x^k} = g^{sum_k s_k x^k}
    (one Fr evaluation + one scalar multiplication on the CPU checks an MSM of any size)."""
    rng = random.Random(14)
    d = 3000
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    g = gpu.SRS.new(d, x, alpha)
    xi = pow(x, -1, R)
    try:
        for N, wb, chunk in ((1, 0, 0), (2, 4, 8), (257, 7, 0), (1000, 0, 0), (4097, 11, 16), (6001, 13, 64), (6001, 16, 0)):
            gpu.set_option("window_bits", wb)
            gpu.set_option("chunk", chunk)
            lo = -(N // 2)
            sc = [rng.randrange(R) for _ in range(N)]
            for fam, mult in ((0, 1), (1, alpha)):
                s = list(sc)
                if fam == 1 and lo <= 0 < lo + N:
                    s[-lo] = 0
                acc = 0
                for k, v in enumerate(s):
                    e = lo + k
                    acc += v * (pow(x, e, R) if e >= 0 else pow(xi, -e, R))
                want = C(bls.g1_mul_gen(acc * mult % R))
                assert gpu.msm(g, fam, lo, s) == want, (N, wb, fam)
    finally:
        gpu.set_option("window_bits", 0)
        gpu.set_option("chunk", 0)


def test_msm_sort_paths_agree(gpu):
    """The tiled counting sort (shared-memory histograms, several tiles per job, one bucket set per
    window or one per job with the precomputed tables) and the first implementation (thread per
    term, global atomics) feed the same buckets: same element, and the trapdoor identity holds."""
    rng = random.Random(21)
    d = 3500
    x = rng.randrange(1, R)
    xi = pow(x, -1, R)
    N = 7000                                  # > 3 tiles of 2048 terms
    pool = [0, 1, R - 1, rng.randrange(R)]
    sc = [rng.randrange(R) if k % 3 else rng.choice(pool) for k in range(N)]
    lo = -(N // 2)
    acc = sum(v * (pow(x, lo + k, R) if lo + k >= 0 else pow(xi, -(lo + k), R)) for k, v in enumerate(sc)) % R
    want = C(bls.g1_mul_gen(acc))
    try:
        for pre in (0, -1):
            gpu.set_option("precompute", pre)
            g = gpu.SRS.new(d, x, 4)
            for wb in ((0, 5, 12, 16) if pre == 0 else (0,)):
                gpu.set_option("window_bits", wb)
                for mode in (0, 1):
                    gpu.set_option("sort_mode", mode)
                    assert gpu.msm(g, 0, lo, sc) == want, (pre, wb, mode)
                    tiles = gpu.last_timing_ms("msm.sort_tiles")
                    assert (tiles >= 4) if mode == 1 else (tiles == 0), (pre, wb, mode, tiles)
            g.free()
    finally:
        gpu.set_option("precompute", -1)
        gpu.set_option("window_bits", 0)
        gpu.set_option("sort_mode", 1)


def test_msm_skewed_and_degenerate(gpu):
    """Zeros, +-1 and repeated scalars (heavy buckets), and x = 1 where every base coincides."""
    rng = random.Random(15)
    for x in (1, rng.randrange(2, R)):
        d = 2500
        g = gpu.SRS.new(d, x, 4)
        N = 5000
        pool = [0, 0, 0, 0, 1, R - 1, 2, rng.randrange(R)]
        s = [rng.choice(pool) for _ in range(N)]
        lo = -d
        xi = pow(x, -1, R)
        acc = sum(v * (pow(x, lo + k, R) if lo + k >= 0 else pow(xi, -(lo + k), R)) for k, v in enumerate(s)) % R
        for wb in (0, 8, 14):
            gpu.set_option("window_bits", wb)
            assert gpu.msm(g, 0, lo, s) == C(bls.g1_mul_gen(acc)), (x, wb)
        gpu.set_option("window_bits", 0)
        # sharded: partial sums of two slices fold to the same element
        parts = [gpu.msm_partial(g, 0, lo, s[:1234]), gpu.msm_partial(g, 0, lo + 1234, s[1234:])]
        assert gpu.g1_sum(parts) == C(bls.g1_mul_gen(acc))


@pytest.mark.parametrize("which", ["example1", "example2"])
def test_prove_reference_circuits_byte_exact(gpu, which):
    rng = random.Random(16)
    circuit, assignment = example1() if which == "example1" else example2(12)
    n, Q = len(assignment.aL), len(circuit.weights.wL)
    gc, ga = to_gpu_types(gpu, circuit, assignment)
    for d in ((12, 25) if n == 1 else (16, 50)):
        x, alpha = rng.randrange(1, R), rng.randrange(1, R)
        g, o = _srs_pair(gpu, d, x, alpha)
        rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
        want, (y, z, yzs) = S.prove(o, assignment, circuit, rnd)
        got = gpu.prove_bytes(g, ga, gc, rnd)
        assert got == S.encode_proof(want)
        assert S.verify_trapdoor(o, circuit, S.decode_proof(got, Q), y, z, yzs)


def test_prove_random_circuits_like_test_sonic(gpu):
    """test/Test/Protocol.hs:14-23 `test_sonic`, on the CUDA path, checked byte for byte against
    the oracle and then by `verify` (trapdoor form)."""
    rng = random.Random(17)
    for trial in range(6):
        circuit, assignment = rnd_circuit(rng, n=rng.randint(1, 12))
        n, Q = len(assignment.aL), len(circuit.weights.wL)
        d = max(random_d(rng, n), 4 * n + 8)
        x, alpha = rng.randrange(1, R), rng.randrange(1, R)
        g, o = _srs_pair(gpu, d, x, alpha)
        gc, ga = to_gpu_types(gpu, circuit, assignment)
        rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
        want, (y, z, yzs) = S.prove_dense(o, assignment, circuit, rnd)
        got = gpu.prove_bytes(g, ga, gc, rnd)
        assert got == S.encode_proof(want), (trial, n, Q, d)
        assert S.verify_trapdoor(o, circuit, S.decode_proof(got, Q), y, z, yzs)


def test_prove_panics_like_the_reference(gpu):
    rng = random.Random(18)
    circuit, assignment = example2(12)
    gc, ga = to_gpu_types(gpu, circuit, assignment)
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(5))]
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    for d in (13, 15):  # 13 < 7n: Protocol.hs:54-55; 15 < 4n+8: commitPoly runs off the SRS
        g, o = _srs_pair(gpu, d, x, alpha)
        with pytest.raises(S.SonicPanic) as oe:
            S.prove(o, assignment, circuit, rnd)
        with pytest.raises(gpu.SonicError) as ge:
            gpu.prove_bytes(g, ga, gc, rnd)
        assert ge.value.text == str(oe.value)
    # an unsatisfied circuit leaves a non-zero X^0 term in t(X,y): commitPoly indexes g^alpha and panics
    g, o = _srs_pair(gpu, 40, x, alpha)
    bad = S.Assignment(list(assignment.aL), list(assignment.aR), [assignment.aO[0] + 1, assignment.aO[1]])
    with pytest.raises(S.SonicPanic) as oe:
        S.prove(o, bad, circuit, rnd)
    with pytest.raises(gpu.SonicError) as ge:
        gpu.prove_bytes(g, gpu.Assignment(bad.aL, bad.aR, bad.aO), gc, rnd)
    assert ge.value.text == str(oe.value)


def test_hsc_prove_vs_oracle(gpu):
    """test/Test/Signature.hs:20-36."""
    rng = random.Random(19)
    for trial in range(3):
        circuit, assignment = rnd_circuit(rng, n=rng.randint(2, 9))
        n, Q = len(assignment.aL), len(circuit.weights.wL)
        d = random_d(rng, n)
        x, alpha = rng.randrange(1, R), rng.randrange(1, R)
        g, o = _srs_pair(gpu, d, x, alpha)
        gc, _ = to_gpu_types(gpu, circuit, assignment)
        m = Q if trial else Q + 2
        yzs = [(rng.randrange(1, R), rng.randrange(1, R)) for _ in range(m)]
        u, v = rng.randrange(1, R), rng.randrange(1, R)
        sXY = S.sPoly(circuit.weights)
        want = S.hscProve(o, sXY, yzs, u, v)
        got = gpu.hscProve(g, gc, yzs, u, v)
        assert [(a, (b, c)) for a, (b, c) in got.hscS] == [(C(a), (b, C(c))) for a, (b, c) in want.hscS]
        assert got.hscW == [(a, C(b), C(c)) for a, b, c in want.hscW]
        assert (got.hscQv, got.hscC, got.hscU, got.hscV) == (C(want.hscQv), C(want.hscC), u, v)
        assert S.hscVerify_trapdoor(o, sXY, yzs, want)


def test_srs_save_load_roundtrip(gpu, tmp_path):
    """a saved SRS loads back to the same resident arrays (same commitments,
    same proof), with and without precomputed levels."""
    rng = random.Random(20)
    circuit, assignment = example2(12)
    gc, ga = to_gpu_types(gpu, circuit, assignment)
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(5))]
    try:
        for pre in (-1, 0):
            gpu.set_option("precompute", pre)
            x, alpha = rng.randrange(1, R), rng.randrange(1, R)
            g = gpu.SRS.new(33, x, alpha)
            path = str(tmp_path / f"srs_{pre}.bin")
            g.save(path)
            g2 = gpu.SRS.load(path)
            assert g2.srsD == 33
            assert g2.gPositiveAlphaX == g.gPositiveAlphaX and g2.gNegativeX == g.gNegativeX
            assert gpu.prove_bytes(g2, ga, gc, rnd) == gpu.prove_bytes(g, ga, gc, rnd)
    finally:
        gpu.set_option("precompute", -1)
    with open(str(tmp_path / "bad.bin"), "wb") as fh:
        fh.write(b"not an srs")
    with pytest.raises(gpu.SonicError):
        gpu.SRS.load(str(tmp_path / "bad.bin"))


def test_sparse_circuit_load_gives_the_same_proof(gpu):
    """the CSR ingestion path builds the same s(X,y), s(u,Y) and therefore
    the same proof bytes as the dense Q x n path (and as the oracle)."""
    rng = random.Random(21)
    cases = [example2(12)] + [rnd_circuit(rng, n=rng.randint(3, 14)) for _ in range(3)]
    # a circuit with genuinely sparse, non-trivial weights (several non-zeros per row and column)
    n, Q = 9, 6
    aL = [rng.randrange(R) for _ in range(n)]
    aR = [rng.randrange(R) for _ in range(n)]
    aO = [a * b % R for a, b in zip(aL, aR)]
    mk = lambda: [[rng.randrange(R) if rng.random() < 0.3 else 0 for _ in range(n)] for _ in range(Q)]
    wL, wR, wO = mk(), mk(), mk()
    dot = lambda v, row: sum(a * b for a, b in zip(v, row))
    cs = [(dot(aL, wL[q]) + dot(aR, wR[q]) + dot(aO, wO[q])) % R for q in range(Q)]
    cases.append((S.ArithCircuit(S.GateWeights(wL, wR, wO), cs), S.Assignment(aL, aR, aO)))
    for circuit, assignment in cases:
        n, Q = len(assignment.aL), len(circuit.weights.wL)
        d = 7 * n + 9
        x, alpha = rng.randrange(1, R), rng.randrange(1, R)
        g, o = _srs_pair(gpu, d, x, alpha)
        rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
        w = circuit.weights
        dense = gpu.ArithCircuit(gpu.GateWeights(w.wL, w.wR, w.wO), circuit.cs)
        sparse = gpu.ArithCircuit(gpu.GateWeights(w.wL, w.wR, w.wO), circuit.cs, sparse=True)
        ga = gpu.Assignment(assignment.aL, assignment.aR, assignment.aO)
        want, _ = S.prove_dense(o, assignment, circuit, rnd)
        a = gpu.prove_bytes(g, ga, dense, rnd)
        b = gpu.prove_bytes(g, ga, sparse, rnd)
        assert a == b == S.encode_proof(want)


def test_pcv_fold_batches_the_verifier_checks(gpu):
    """the 3Q+4 pcV checks of a proof folded into one multi-pairing's G1
    inputs.  Checked in the exponent with the trapdoor:
        alpha*x*A + alpha*B == sum_m x^(-d+max_m) * C_m   holds for a valid proof, fails for a forged one."""
    rng = random.Random(22)
    circuit, assignment = example2(12)
    n, Q, d = 2, 5, 20
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    g, o = _srs_pair(gpu, d, x, alpha)
    gc, ga = to_gpu_types(gpu, circuit, assignment)
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
    proof, orc = gpu.prove(g, ga, gc, rnd)
    y, z, yzs = orc.rndOracleY, orc.rndOracleZ, orc.rndOracleYZs
    ky = sum(k * S.fr_pow(y, n + 1 + q) for q, k in enumerate(circuit.cs)) % R
    t = (proof.prA * (proof.prB + proof.prS) - ky) % R
    h = proof.prHscProof
    sv = S.dense_sXy(circuit.weights, h.hscV).eval(h.hscU)
    # group 0: max = n (commitment R), group 1: max = d (everything else) -- Protocol.hs:121-124, Signature.hs:80-90
    checks = [(proof.prR, z, (proof.prA, proof.prWa), 0), (proof.prR, y * z % R, (proof.prB, proof.prWb), 0),
              (proof.prT, z, (t, proof.prWt), 1), (h.hscC, h.hscV, (sv, h.hscQv), 1)]
    for (yi, zi), (ci, (si, wi)), (sip, wip, qi) in zip(yzs, h.hscS, h.hscW):
        checks += [(ci, zi, (si, wi), 1), (ci, h.hscU, (sip, wip), 1), (h.hscC, yi, (sip, qi), 1)]
    assert len(checks) == 3 * Q + 4
    weights = [rng.randrange(1, R) for _ in checks]
    D = bls.g1_decompress

    def holds(A, B, Cs):
        lhs = bls.g1_add(bls.g1_mul(D(A), alpha * x % R), bls.g1_mul(D(B), alpha))
        rhs = bls.g1_add(bls.g1_mul(D(Cs[0]), S.fr_pow(x, -d + n)), bls.g1_mul(D(Cs[1]), S.fr_pow(x, 0)))
        return lhs == rhs

    A, B, Cs = gpu.pcv_fold(checks, weights)
    assert holds(A, B, Cs)
    # the fold equals the straightforward sums
    accA = bls.INF
    for (F, zi, (vi, Wi), grp), r in zip(checks, weights):
        accA = bls.g1_add(accA, bls.g1_mul(D(Wi), r))
    assert A == C(accA)
    forged = list(checks)
    F0, z0, (v0, W0), g0 = forged[2]
    forged[2] = (F0, z0, ((v0 + 1) % R, W0), g0)
    assert not holds(*gpu.pcv_fold(forged, weights))
    with pytest.raises(gpu.SonicError):
        gpu.pcv_fold([(bytes(48), 1, (1, bytes(48)), 0)], [1])


def test_prove_with_more_than_64_msms(gpu):
    """Q = 17 gives 4Q+7 = 75 MSMs: more than one batch of the MSM pipeline (64 jobs per launch)."""
    rng = random.Random(23)
    circuit, assignment = rnd_circuit(rng, n=18, m=17)
    n, Q = 18, 17
    d = 7 * n + 3
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    g, o = _srs_pair(gpu, d, x, alpha)
    gc, ga = to_gpu_types(gpu, circuit, assignment)
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
    want, (y, z, yzs) = S.prove_dense(o, assignment, circuit, rnd)
    got = gpu.prove_bytes(g, ga, gc, rnd)
    assert got == S.encode_proof(want)
    # the same proof assembled from two shards (both sharding modes are functions of the sizes only)
    blobs = [gpu.prove_shard(g, ga, gc, rnd, r, 2) for r in range(2)]
    assert gpu.prove_combine(Q, blobs) == got
    blobs = [gpu.prove_shard(g, ga, gc, rnd, r, 3) for r in range(3)]
    assert gpu.prove_combine(Q, blobs) == got
    # more ranks than some MSMs have terms, and a rank count that does not divide anything
    for world in (5, 8):
        blobs = [gpu.prove_shard(g, ga, gc, rnd, r, world) for r in range(world)]
        assert gpu.prove_combine(Q, blobs) == got, world


def test_srs_g2_vectors(gpu):
    """the h-vectors of SRS.new (SRS.hs:35-36,40-41) from the G2 fixed-base batch."""
    rng = random.Random(24)
    d = 11
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    try:
        gpu.set_option("g2", 1)
        g = gpu.SRS.new(d, x, alpha)
    finally:
        gpu.set_option("g2", 0)
    xi = pow(x, -1, R)
    H = bls.G2_GEN
    c2 = bls.g2_compress
    assert g.hPositiveX == [c2(bls.g2_mul(H, pow(x, i, R))) for i in range(0, d + 1)]
    assert g.hNegativeX == [c2(bls.g2_mul(H, pow(xi, i, R))) for i in range(1, d + 1)]
    assert g.hPositiveAlphaX == [c2(bls.g2_mul(H, alpha * pow(x, i, R) % R)) for i in range(0, d + 1)]
    assert g.hNegativeAlphaX == [c2(bls.g2_mul(H, alpha * pow(xi, i, R) % R)) for i in range(1, d + 1)]
    g1only = gpu.SRS.new(d, x, alpha)
    with pytest.raises(gpu.SonicError):
        g1only.hPositiveX


def test_prove_degenerate_inputs(gpu):
    """Edge cases: Q = 1; an all-zero assignment (r'(X,1) is the four blinders only, most MSM scalars
    are zero); zero blinders; hscProve with no (y, z) pairs at all."""
    rng = random.Random(25)
    x, alpha = rng.randrange(1, R), rng.randrange(1, R)
    # Q = 1, n = 3
    circuit, assignment = rnd_circuit(rng, n=3, m=1)
    g, o = _srs_pair(gpu, 25, x, alpha)
    gc, ga = to_gpu_types(gpu, circuit, assignment)
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(1))]
    assert gpu.prove_bytes(g, ga, gc, rnd) == S.encode_proof(S.prove_dense(o, assignment, circuit, rnd)[0])
    # all-zero assignment satisfies any circuit with cs = 0; blinders zero as well
    n, Q = 4, 3
    w = [[rng.randrange(R) for _ in range(n)] for _ in range(Q)]
    circuit = S.ArithCircuit(S.GateWeights(w, [list(r) for r in w], [list(r) for r in w]), [0] * Q)
    assignment = S.Assignment([0] * n, [0] * n, [0] * n)
    g, o = _srs_pair(gpu, 7 * n + 5, x, alpha)
    gc, ga = to_gpu_types(gpu, circuit, assignment)
    for blind in (True, False):
        rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
        if not blind:
            rnd[0:4] = [0, 0, 0, 0]
        want, (y, z, yzs) = S.prove_dense(o, assignment, circuit, rnd)
        got = gpu.prove_bytes(g, ga, gc, rnd)
        assert got == S.encode_proof(want)
        assert S.verify_trapdoor(o, circuit, S.decode_proof(got, Q), y, z, yzs)
        if not blind:
            assert want.prR is bls.INF  # r'(X,1) = 0: the commitment is the identity
    # a zero challenge is `recip 0` (every challenge evaluates a polynomial with negative powers)
    rnd[5] = 0
    with pytest.raises(gpu.SonicError) as e:
        gpu.prove_bytes(g, ga, gc, rnd)
    assert e.value.kind == "DIV_BY_ZERO"
    # hscProve with an empty list of pairs
    u, v = rng.randrange(1, R), rng.randrange(1, R)
    sXY = S.sPoly(circuit.weights)
    want = S.hscProve(o, sXY, [], u, v)
    got = gpu.hscProve(g, gc, [], u, v)
    assert (got.hscS, got.hscW, got.hscQv, got.hscC) == ([], [], C(want.hscQv), C(want.hscC))
