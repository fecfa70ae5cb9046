"""Multi-GPU plumbing: one process per GPU, torch.distributed for the exchange.

The path shards two ways: across independent proofs (no exchange at all)
and, for one proof, across equal runs of the proof's MSM terms (`deal_terms`; a standalone MSM is
cut into contiguous slices, `slice_bounds`).  The second has exactly one
exchange step: each rank ends with 4Q+7 partial G1 sums (96 B each) plus the field values;
one all-gather of that small blob (NCCL over NVLink on GPUs, gloo in the CPU tests) and a
fold with `<>` complete the proof.  Nothing else crosses ranks: the SRS is replicated.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def slice_bounds(lo: int, hi: int, rank: int, world: int) -> Tuple[int, int]:
    """The contiguous part [a, b) of an exponent window [lo, hi) that `rank` of `world` sums in a
    standalone sharded MSM."""
    span = hi - lo
    return lo + span * rank // world, lo + span * (rank + 1) // world


def deal_terms(lengths: Sequence[int], world: int, t_msms: Sequence[int] = (), t_extra: int = 0) -> List[List[Tuple[int, int]]]:
    """How prove_run (csrc/prove.cu) shards one proof: the (clipped) exponent windows of all MSMs,
    concatenated in record order, are cut into `world` runs of terms.  Returns, per rank, the
    part [a, b) of every MSM's window (offsets into the window; a == b where the rank holds nothing).
    A rank owns a few whole MSMs and at most two partial ones; at most world-1 MSMs are split; a boundary
    never leaves a sliver of an MSM (fewer than min(len/2, max(4096, total/(64 world))) terms) on either side.
    The runs are equal, except that ranks owning part of the MSMs `t_msms` (prT and prWt, records 1
    and 4 of `prove`: those ranks also build t(X,y)) are dealt `t_extra` (= n/2 * min(world - 2, 6) / 6) terms less when
    that leaves the same ranks in charge of them."""
    total = sum(lengths)
    pos = [0]
    for n in lengths:
        pos.append(pos[-1] + n)

    def owners(bd):
        return [any(min(bd[r + 1], pos[i + 1]) > max(bd[r], pos[i]) for i in t_msms) for r in range(world)]

    def snap(bd):
        """No slivers: a boundary that would leave only a few terms of an MSM on one side moves to that MSM's border."""
        floor_ = max(4096, total // (world * 64))
        for r in range(1, world):
            b = bd[r]
            for i, n in enumerate(lengths):
                if pos[i] < b < pos[i + 1]:
                    minp = min(n // 2, floor_)
                    if b - pos[i] < minp:
                        bd[r] = pos[i]
                    elif pos[i + 1] - b < minp:
                        bd[r] = pos[i + 1]
                    break
        return bd

    bound = snap([total * r // world for r in range(world + 1)])
    if t_msms and t_extra > 0:
        o1 = owners(bound)
        k = sum(o1)
        padded = total + k * t_extra
        b2, ok = [0], 0 < k < world
        for r in range(world):
            if not ok:
                break
            cap = padded * (r + 1) // world - padded * r // world
            if o1[r] and cap < t_extra:
                ok = False
                break
            b2.append(b2[-1] + cap - (t_extra if o1[r] else 0))
        if ok and b2[-1] == total and owners(snap(b2)) == o1:
            bound = b2
    out = []
    for rank in range(world):
        lo, hi = bound[rank], bound[rank + 1]
        out.append([(min(max(lo - p, 0), n), min(max(hi - p, 0), n)) for p, n in zip(pos, lengths)])
    return out


def value_leads(deal: List[List[Tuple[int, int]]], Q: int) -> List[int]:
    """Which rank contributes each device-computed field value of a sharded proof (mirror of prove.cu): the lowest
    rank holding terms of the MSM the value's opening belongs to (rank 0 for an empty window).  Values in record
    order: prA (opening prWa, record 2), prB (prWb, 3), prS (led by prT's first rank, 1), s_j (W_j), s'_j (Q_j).
    Every other rank leaves zeros in its exchange record and the fold ORs the bytes."""
    def first(i):
        for r, parts in enumerate(deal):
            if parts[i][1] > parts[i][0]:
                return r
        return 0

    base = 5
    msm_of = [2, 3, 1] + [base + 2 * j + 1 for j in range(Q)] + [base + 2 * Q + 2 * j + 1 for j in range(Q)]
    return [first(i) for i in msm_of]


def all_gather_bytes(blob: bytes, group=None) -> List[bytes]:
    """All-gather of one fixed-size byte blob per rank; returns the blobs in rank order."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    flat = out.cpu().numpy().tobytes()
    n = len(blob)
    return [flat[i * n:(i + 1) * n] for i in range(world)]


def prove_sharded(srs, assignment, circuit, rnd: Sequence[int], group=None) -> bytes:
    """One proof over all ranks of `group`; every rank returns the same proof bytes."""
    from . import api

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return api.prove_bytes(srs, assignment, circuit, rnd)
    blob = api.prove_shard(srs, assignment, circuit, rnd, rank, world)
    blobs = all_gather_bytes(blob, group)
    return api.prove_combine(circuit.handle().Q, blobs)
