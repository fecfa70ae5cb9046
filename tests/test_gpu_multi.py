"""GPU: the C consumer of the ABI (tests/c/abi_multi.c, no Python in the process) binds 1 device and then every
visible device IN ONE PROCESS and demands byte-identical SRS elements, proofs, batches, commitments, openings, MSMs and
panic texts.  On a one-GPU box this exercises the single-device path from plain C; on a multi-GPU box the sharded
SRS.new + NCCL all-gather, the in-library sharded prove and the batch dealing."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(ndev, log_n):
    cdir = os.path.join(ROOT, "tests", "c")
    subprocess.run(["make", "-C", cdir, "abi_multi"], check=True, capture_output=True, text=True)
    out = subprocess.run([os.path.join(cdir, "abi_multi"), str(ndev), str(log_n)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr + out.stdout
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_c_consumer_one_process_all_devices(gpu):
    import torch

    ndev = min(torch.cuda.device_count(), 8)
    for log_n in (6, 12):
        line = _run(ndev, log_n)
        assert line["identical"] is True and line["ndev"] == ndev
        assert line["panic"] == "commitPoly: gNegativeAlphaX is not long enough: -1 >= %d" % line["d"]


def test_python_mirror_in_library_sharding(gpu):
    """The same through the ctypes mirror in a fresh process (this process is already bound to one device): proof bytes from
    sonic_init over all visible devices equal the one-device proof and the oracle's."""
    import sys

    import torch

    ndev = min(torch.cuda.device_count(), 8)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_multi.py"), str(ndev)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok %d" % ndev), out.stderr + out.stdout
