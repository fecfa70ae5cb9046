#!/usr/bin/env python3
"""Per-call timing of a sharded prove() under torchrun (diagnostic): resident vs host-buffer inputs.

torchrun --nproc-per-node 2 tools/diag_shard.py
"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import sonic_b200 as sb
from sonic_b200 import capi, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sb.init(local)
L = capi.lib()
n, Q = 1 << 16, 8
x, alpha = synth.trapdoor()
srs = sb.SRS.new(7 * n, x, alpha)
c = synth.synthetic_circuit_bytes(n, Q, seed=4)
ch = ctypes.c_void_p()
capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
nr = 2 * Q + 8
host_in = torch.empty(3 * n * 32, dtype=torch.uint8).pin_memory()
host_in.numpy()[:] = np.concatenate([c["aL"], c["aR"], c["aO"]])
host_rnd = torch.empty(nr * 32, dtype=torch.uint8).pin_memory()
host_rnd.numpy()[:] = np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(40, nr)]), dtype=np.uint8)
hin, hrnd = host_in.data_ptr(), host_rnd.data_ptr()
d_in, d_rnd = ctypes.c_void_p(), ctypes.c_void_p()
capi.check(L.sonic_dev_alloc(3 * n * 32, ctypes.byref(d_in)))
capi.check(L.sonic_dev_alloc(nr * 32, ctypes.byref(d_rnd)))
capi.check(L.sonic_dev_upload(d_in, hin, 3 * n * 32))
capi.check(L.sonic_dev_upload(d_rnd, hrnd, nr * 32))
nm = 4 * Q + 7
out = ctypes.create_string_buffer(max(int(L.sonic_proof_size(Q)), int(L.sonic_shard_blob_size(Q))))
w = ctypes.c_uint64(0)
part_t = torch.empty(nm * 96, dtype=torch.uint8, device="cuda")


def call(resident: bool):
    t0 = time.perf_counter()
    if resident:
        capi.check(L.sonic_prove_shard_sink(srs._h, ch, d_in, 1, d_rnd, hrnd, rank, world, out, len(out), ctypes.byref(w), part_t.data_ptr()))
    else:
        capi.check(L.sonic_prove_shard_sink(srs._h, ch, hin, 0, None, hrnd, rank, world, out, len(out), ctypes.byref(w), part_t.data_ptr()))
    dt = 1e3 * (time.perf_counter() - t0)
    return dt, {k: round(sb.last_timing_ms(k), 2) for k in ("total", "poly", "msm.sort", "msm.accumulate", "msm.reduce", "msm.terms")}


for label, seq in (("warm", [True, False, True, False]), ("resident", [True] * 4), ("host", [False] * 4), ("resident", [True] * 4), ("alternate", [True, False] * 3)):
    dist.barrier()
    torch.cuda.synchronize()
    for r in seq:
        dt, st = call(r)
        print("rank %d %-9s %s wall %.2f ms  %s" % (rank, label, "res " if r else "host", dt, st), flush=True)
dist.destroy_process_group()
