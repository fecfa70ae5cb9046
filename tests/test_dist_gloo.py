"""CPU, world_size 2, gloo: the multi-GPU path's host logic -- dealing equal runs of the proof's
MSM terms to the ranks, the all-gather of the shard blobs (sonic_b200/dist.py) and the fold -- with the oracle
standing in for the CUDA kernels.  The sharded proof must equal the single-rank proof."""
import os
import random
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import bls12_381 as bls
    from oracle import sonic as S
    from sonic_b200 import dist as sdist
    from tests.util import example2

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(77)  # same inputs on every rank
    R = bls.R
    circuit, assignment = example2(12)
    Q, d = 5, 21
    srs = S.srs_new(d, rng.randrange(1, R), rng.randrange(1, R))
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(Q))]
    msms, fvals = S.prove_dense_plan(srs, assignment, circuit, rnd)
    assert len(msms) == 4 * Q + 7 and len(fvals) == 2 * Q + 5
    # this rank's shard blob: raw partial sums over its slice of every MSM, then the Fr values
    parts = []
    windows = [(max(lo, -d), min(lo + len(scal), d + 1)) for _, lo, scal in msms]
    deal = sdist.deal_terms([chi - clo for clo, chi in windows], world, t_msms=(1, 4), t_extra=12 // 2)
    mine = deal[rank]
    for (alpha, lo, scal), (clo, chi), (a, b) in zip(msms, windows, mine):
        parts.append(bls.g1_to_raw(S.fold_msm(srs, (alpha, lo, scal), clo + a, clo + b)))
    # the exchange record of this rank (include/sonic_b200.h: sonic_prove_shard): raw partial sums, then the field values
    # this rank LEADS (zeros elsewhere), then hscU, hscV which every rank knows
    leads = sdist.value_leads(deal, Q)
    nv = 2 * Q + 3
    vals = [bls.fr_to_bytes(v) if leads[k] == rank else bytes(32) for k, v in enumerate(fvals[:nv])]
    blob = b"".join(parts) + b"".join(vals) + b"".join(bls.fr_to_bytes(v) for v in fvals[nv:])
    blobs = sdist.all_gather_bytes(blob)
    assert len(blobs) == world and blobs[rank] == blob
    nm = len(msms)
    g48 = []
    for m in range(nm):
        acc = bls.INF
        for r in range(world):
            acc = bls.g1_add(acc, bls.g1_from_raw(blobs[r][96 * m:96 * m + 96]))
        g48.append(bls.g1_compress(acc))
    # fold of the values: exactly one rank contributed each, so the bytes are OR-ed
    fv = []
    for i in range(len(fvals)):
        if i < nv:
            word = 0
            for r in range(world):
                word |= int.from_bytes(blobs[r][96 * nm + 32 * i:96 * nm + 32 * i + 32], "little")
            fv.append(word)
        else:
            fv.append(bls.fr_from_bytes(blobs[rank][96 * nm + 32 * i:96 * nm + 32 * i + 32]))
    assert sorted(set(leads)) == sorted(set(leads) & set(range(world))) and len(leads) == nv
    proof = S.assemble_proof_bytes(Q, g48, fv)
    want, _ = S.prove_dense(srs, assignment, circuit, rnd)
    ok = proof == S.encode_proof(want)
    # the slices tile every window exactly
    for lo, hi in ((-7, 9), (0, 1), (-100, 101), (5, 5)):
        cuts = [sdist.slice_bounds(lo, hi, r, world) for r in range(world)]
        ok = ok and cuts[0][0] == lo and cuts[-1][1] == hi and all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
    with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as fh:
        fh.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_proof_equals_single_rank_proof(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"


def test_plan_matches_prove_dense():
    sys.path.insert(0, ROOT)
    from oracle import bls12_381 as bls
    from oracle import sonic as S
    from tests.util import example1

    rng = random.Random(78)
    R = bls.R
    circuit, assignment = example1()
    srs = S.srs_new(13, rng.randrange(1, R), rng.randrange(1, R))
    rnd = [rng.randrange(1, R) for _ in range(S.rnd_count(2))]
    msms, fvals = S.prove_dense_plan(srs, assignment, circuit, rnd)
    g48 = [bls.g1_compress(S.fold_msm(srs, m)) for m in msms]
    want, _ = S.prove_dense(srs, assignment, circuit, rnd)
    assert S.assemble_proof_bytes(2, g48, fvals) == S.encode_proof(want)


def test_term_dealing_is_balanced_and_total():
    sys.path.insert(0, ROOT)
    from sonic_b200 import dist as sdist

    n, Q = 1 << 16, 8
    r, t, s_, c = 3 * n + 5, 7 * n + 9, 3 * n + 1, 2 * n + Q + 1
    lengths = [r, t, r - 1, r - 1, t - 1] + [s_, s_ - 1] * Q + [s_ - 1, c - 1] * Q + [c - 1, c]
    assert len(lengths) == 4 * Q + 7
    for lens in (lengths, [1000, 10, 10], [5], [0, 3, 0, 0, 4], [1] * 7):
        for world in (2, 3, 4, 8):
            deal = sdist.deal_terms(lens, world)
            load = [sum(b - a for a, b in parts) for parts in deal]
            # equal runs, except that a boundary never leaves a sliver of an MSM on either side (it snaps to the border)
            floor_ = max(4096, sum(lens) // (world * 64))
            assert sum(load) == sum(lens) and max(load) - min(load) <= 1 + 2 * min(floor_, max(lens))
            for parts in deal:
                for (a, b), L_ in zip(parts, lens):
                    assert b == a or b - a == L_ or (a > 0 and b < L_) or b - a >= min(L_ // 2, floor_), (a, b, L_)   # a piece touching neither border is a whole run
            if len(lens) > 4:
                # ranks that also build t(X,y) (owners of records 1 and 4) are dealt t_extra terms less
                extra = n // 2 if lens is lengths else 1
                deal = sdist.deal_terms(lens, world, t_msms=(1, 4), t_extra=extra)
                load2 = [sum(b - a for a, b in parts) for parts in deal]
                own = [any(parts[i][1] > parts[i][0] for i in (1, 4)) for parts in deal]
                assert sum(load2) == sum(lens)
                if load2 != load:
                    heavy = [l for l, o in zip(load2, own) if o]
                    light = [l for l, o in zip(load2, own) if not o]
                    tol = 1 + 2 * min(floor_, max(lens))
                    assert heavy and light and max(heavy) - min(heavy) <= tol and max(light) - min(light) <= tol
                    assert abs((min(light) - max(heavy)) - extra) <= 2 * tol
            split = 0
            for m, L in enumerate(lens):
                cuts = [deal[k][m] for k in range(world) if deal[k][m][1] > deal[k][m][0]]
                # the parts of one MSM tile its window exactly, in rank order
                assert sum(b - a for a, b in cuts) == L
                assert all(cuts[i][1] == cuts[i + 1][0] for i in range(len(cuts) - 1))
                assert not cuts or (cuts[0][0] == 0 and cuts[-1][1] == L)
                split += len(cuts) > 1
            assert split <= world - 1
            # a rank's run is contiguous: at most two of its MSMs are partial
            assert all(sum(1 for (a, b), L in zip(parts, lens) if 0 < b - a < L) <= 2 for parts in deal)
