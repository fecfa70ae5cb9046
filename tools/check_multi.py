#!/usr/bin/env python3
"""In-library multi-GPU check through the ctypes mirror (fresh process): SRS vectors, proof and batch from
sonic_init over `ndev` devices against the Python oracle.  tests/test_gpu_multi.py runs it."""
import faulthandler
import os
import random
import sys

faulthandler.dump_traceback_later(int(os.environ.get("SONIC_CHECK_TRACE_AFTER", "100")), exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("SONIC_CHECK_IMPORT_TORCH"):
    import torch  # noqa: F401
import sonic_b200 as sb  # noqa: E402
from oracle import bls12_381 as bls, sonic as S  # noqa: E402
from tests.util import rnd_circuit, to_gpu_types  # noqa: E402

ndev = int(sys.argv[1])
say = lambda *a: print(*a, flush=True)
sb.init(list(range(ndev)))
assert sb.device_count() == ndev
say("init ok")
rng = random.Random(5)
R = bls.R
circuit, assignment = rnd_circuit(rng, 11, 4)
d = 90
x, alpha = rng.randrange(1, R), rng.randrange(1, R)
rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(4))]
g = sb.SRS.new(d, x, alpha)
say("srs ok")
o = S.srs_new(d, x, alpha)
assert g.gPositiveX == [bls.g1_compress(p) for p in o.gPositiveX]
assert g.gNegativeAlphaX == [bls.g1_compress(p) for p in o.gNegativeAlphaX]
say("srs vectors ok")
gc, ga = to_gpu_types(sb, circuit, assignment)
want = S.encode_proof(S.prove_dense(o, assignment, circuit, rnd)[0])
say("oracle ok")
assert sb.prove_bytes(g, ga, gc, rnd) == want
say("prove ok")
assert sb.prove_batch(g, [ga] * 5, gc, [rnd] * 5) == [want] * 5
# persistence with several devices: the file is read once and uploaded to every device
import tempfile
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "srs.bin")
    g.save(path)
    g2 = sb.SRS.load(path)
    assert g2.srsD == d and g2.gNegativeX == g.gNegativeX
    assert sb.prove_bytes(g2, ga, gc, rnd) == want
say("save/load ok")
# a standalone MSM cut across the devices, and an unsatisfied assignment: the same panic as on one device
sb.set_option("shard_min_terms", 8)
sc = [rng.randrange(R) for _ in range(2 * d + 1)]
xi = pow(x, -1, R)
acc = sum(v * (pow(x, k - d, R) if k >= d else pow(xi, d - k, R)) for k, v in enumerate(sc)) % R
assert sb.msm(g, 0, -d, sc) == bls.g1_compress(bls.g1_mul_gen(acc))
bad = sb.Assignment(list(assignment.aL), list(assignment.aR), [(assignment.aO[0] + 1) % R] + list(assignment.aO[1:]))
try:
    sb.prove_bytes(g, bad, gc, rnd)
    raise SystemExit("an unsatisfied assignment produced a proof")
except sb.SonicError as e:
    assert e.text == "commitPoly: gNegativeAlphaX is not long enough: -1 >= %d" % d, e.text
say("msm + panic ok")
print("ok", ndev, flush=True)
