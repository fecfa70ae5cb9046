#!/usr/bin/env python3
"""Workload for compute-sanitizer (memcheck / racecheck / initcheck): SRS.new, a small prove() through every
kernel of the path (tiled sort, compact accumulate, fix-up, heavy buckets, reduction, NTT, open), one
commit/open pair and one standalone MSM; results are compared with the oracle so that a sanitizer-clean
run is also a correct one."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sonic_b200 as sb  # noqa: E402
from oracle import bls12_381 as bls  # noqa: E402
from oracle import sonic as S  # noqa: E402
from tests.util import rnd_circuit, to_gpu_types  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
sb.init(0)
rng = random.Random(7)
R = bls.R
n = 1 << log_n
circuit, assignment = rnd_circuit(rng, n, 3)
d = 7 * n + 3
x, alpha = rng.randrange(1, R), rng.randrange(1, R)
rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(3))]
srs_g = sb.SRS.new(d, x, alpha)
gc, ga = to_gpu_types(sb, circuit, assignment)
got = sb.prove_bytes(srs_g, ga, gc, rnd)
srs_o = S.srs_new(d, x, alpha)
want, _ = S.prove_dense(srs_o, assignment, circuit, rnd)
assert got == S.encode_proof(want)
f = {e: rng.randrange(R) for e in range(-2 * n, 2 * n + 1) if e != 0}
assert sb.commitPoly(srs_g, d, f) == bls.g1_compress(S.commitPoly(srs_o, d, f))
v, w = sb.openPoly(srs_g, rnd[5], f)
vo, wo = S.openPoly(srs_o, rnd[5], f)
assert v == vo and w == bls.g1_compress(wo)
print("sanitize workload ok: n =", n, "launches", sb.launch_count(), flush=True)
