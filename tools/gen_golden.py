#!/usr/bin/env python3
"""Generates tests/golden/*.json from the CPU oracle (seeded).  The fixtures freeze the
oracle's outputs; the CUDA path is tested against them on the GPU box.  The reference itself
cannot produce them (no GHC in this environment)."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bls12_381 as bls  # noqa: E402
from oracle import sonic as S  # noqa: E402
from tests.util import example1, example2, rnd_circuit  # noqa: E402

R = bls.R


def main():
    rng = random.Random(0x534F4E4943)
    out = []
    specs = [("arithCircuitExample1 d=12", example1(), 12), ("arithCircuitExample1 d=25 x=1 alpha=4 (bench/Main.hs)", example1(), 25),
             ("arithCircuitExample2 d=16", example2(12), 16), ("arithCircuitExample2 d=50 x=1 alpha=4 (bench/Main.hs)", example2(12), 50),
             ("rndCircuit n=5 Q=3", rnd_circuit(rng, 5, 3), 47), ("rndCircuit n=9 Q=4", rnd_circuit(rng, 9, 4), 80)]
    for name, (circuit, assignment), d in specs:
        if "x=1" in name:
            x, alpha = 1, 4
        else:
            x, alpha = rng.randrange(1, R), rng.randrange(1, R)
        Q = len(circuit.weights.wL)
        rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
        srs = S.srs_new(d, x, alpha)
        proof, (y, z, yzs) = S.prove_dense(srs, assignment, circuit, rnd)
        assert S.verify_trapdoor(srs, circuit, proof, y, z, yzs)
        w = circuit.weights
        out.append(dict(name=name, d=d, x=x, alpha=alpha, wL=w.wL, wR=w.wR, wO=w.wO, cs=circuit.cs,
                        aL=assignment.aL, aR=assignment.aR, aO=assignment.aO, rnd=rnd,
                        proof_hex=S.encode_proof(proof).hex()))
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "prove_small.json"), "w") as fh:
        json.dump(out, fh)
    co = []
    for d, maxm, lo, hi in ((20, 20, -20, 20), (33, 12, -12, 12), (64, 64, -40, 64)):
        x, alpha, z = rng.randrange(1, R), rng.randrange(1, R), rng.randrange(1, R)
        f = {e: rng.choice([1, R - 1, rng.randrange(R)]) for e in range(lo, hi + 1) if e != -(d - maxm) and rng.random() < 0.85}
        srs = S.srs_new(d, x, alpha)
        v, w = S.openPoly(srs, z, f)
        co.append(dict(d=d, x=x, alpha=alpha, z=z, max=maxm, f={str(e): c for e, c in f.items()},
                       commit_hex=bls.g1_compress(S.commitPoly(srs, maxm, f)).hex(), value=v,
                       open_hex=bls.g1_compress(w).hex()))
    with open(os.path.join(ROOT, "tests", "golden", "commit_open.json"), "w") as fh:
        json.dump(co, fh)
    print("wrote", len(out), "proofs and", len(co), "commit/open cases")


if __name__ == "__main__":
    main()
