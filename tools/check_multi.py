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
print("ok", ndev, flush=True)
