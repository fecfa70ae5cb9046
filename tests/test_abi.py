"""CPU: the C-ABI library loads and exports every symbol include/sonic_b200.h declares; the
host-side logic (argument checks, error text) behaves without a GPU; there is no CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sonic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sonic_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sonic_b200 import capi

    names = _declared_symbols()
    assert len(names) >= 25
    L = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/sonic_b200.h but not exported"
    # the Python binding covers the whole header too
    assert set(names) == set(capi.SYMBOLS), set(names) ^ set(capi.SYMBOLS)


def test_no_cpu_fallback_without_device():
    """Without a usable GPU every compute entry point fails loudly instead of computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    from sonic_b200 import capi

    L = capi.lib()
    dev = (ctypes.c_int * 1)(0)
    rc = L.sonic_init(dev, 1)
    assert rc == 8, rc  # SONIC_ERR_NO_DEVICE
    assert "no CPU fallback" in capi.last_error()
    h = ctypes.c_void_p()
    one = (1).to_bytes(32, "little")
    assert L.sonic_srs_new(4, one, one, ctypes.byref(h)) == 9  # NOT_INITIALISED
    out = ctypes.create_string_buffer(48)
    assert L.sonic_g1_sum(None, 0, out) == 9
    import sonic_b200

    with pytest.raises(sonic_b200.SonicError) as e:
        sonic_b200.SRS.new(4, 3, 5)
    assert e.value.kind == "NO_DEVICE"


def test_host_side_argument_checks():
    from sonic_b200 import capi

    L = capi.lib()
    h = ctypes.c_void_p()
    one = (1).to_bytes(32, "little")
    zero = bytes(32)
    r = (0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001).to_bytes(32, "little")
    assert L.sonic_srs_new(4, zero, one, ctypes.byref(h)) == 4       # recip 0
    assert "recip 0" in capi.last_error()
    assert L.sonic_srs_new(4, r, one, ctypes.byref(h)) == 5          # x = r is not canonical
    assert L.sonic_srs_new(0, one, one, ctypes.byref(h)) == 1
    assert L.sonic_rnd_count(8) == 24
    assert L.sonic_proof_size(8) == 39 * 48 + 21 * 32
    assert L.sonic_strerror(3) == b"parameter d is not large enough"
    assert L.sonic_set_option(b"window_bits", 99) == 1
    assert L.sonic_set_option(b"window_bits", 0) == 0


def test_synthetic_generators_are_deterministic_and_canonical():
    from sonic_b200 import synth

    a = synth.fr_bytes_fast(7, 1000)
    b = synth.fr_bytes_fast(7, 1000)
    assert (a == b).all() and a.shape == (1000, 32)
    R = synth.R_MODULUS
    for row in a[:50]:
        assert int.from_bytes(bytes(row), "little") < R
    sk = synth.skewed_fr_bytes(3, 4000)
    vals = [int.from_bytes(bytes(r), "little") for r in sk]
    assert 0.4 < sum(v == 0 for v in vals) / 4000 < 0.6
    assert 0.15 < sum(v in (1, R - 1) for v in vals) / 4000 < 0.35
    (wL, wR, wO, cs), (aL, aR, aO) = synth.synthetic_circuit(6, 3, 2)
    for q in range(3):
        dot = lambda v, row: sum(x * y for x, y in zip(v, row))
        assert (dot(aL, wL[q]) + dot(aR, wR[q]) + dot(aO, wO[q])) % R == cs[q]
    c = synth.synthetic_circuit_bytes(6, 3, 2)
    assert c["ints"]["cs"] == cs and bytes(c["aL"]) == synth.ints_to_bytes(aL)


def test_product_path_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under sonic_b200/ (Python or CUDA) may import,
    include, link or execute it, and the CUDA sources must not read the reference tree either."""
    bad = []
    for root, _dirs, files in os.walk(os.path.join(ROOT, "sonic_b200")):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(root, f), errors="replace").read()
            for needle in ("import oracle", "from oracle", "oracle/", "libsonic_oracle", "/root/reference"):
                if needle in text:
                    bad.append((f, needle))
    assert not bad, bad
    # the boundary header carries no torch / C++ types
    hdr = open(os.path.join(ROOT, "include", "sonic_b200.h")).read()
    assert "torch" not in hdr and "std::" not in hdr and "template" not in hdr


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores) emits one JSON
    line with the keys the driver reads; rank != 0 under torchrun stays silent."""
    import json
    import subprocess
    import sys

    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpoints/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
    env["RANK"] = "1"
    env["WORLD_SIZE"] = "2"
    quiet = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, env=env, timeout=300)
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_python_mirror_checks_shapes_before_any_pointer_crosses():
    """The C side copies Q*n (and Q, n, 2Q+8) field elements from the buffers it is handed: ragged weight rows, a short
    cs, a short assignment or a short list of draws must be refused by the mirror itself (no GPU needed to see it)."""
    import sonic_b200 as sb

    ok = [[1, 2, 3], [4, 5, 6]]
    for wL, wR, wO, cs in (([[1, 2, 3], [4, 5]], ok, ok, [1, 2]),        # ragged row
                           (ok, [[1, 2, 3]], ok, [1, 2]),                 # wR has fewer rows
                           (ok, ok, [[1, 2], [3, 4]], [1, 2]),            # wO is narrower
                           (ok, ok, ok, [1])):                            # cs too short
        with pytest.raises(sb.SonicError) as e:
            sb.ArithCircuit(sb.GateWeights(wL, wR, wO), cs).handle()
        assert e.value.kind == "INVALID_ARG"
    with pytest.raises(sb.SonicError) as e:
        sb.ArithCircuit(sb.GateWeights([], [], []), []).handle()
    assert e.value.text == "Empty weights"

    class FakeHandle:
        n, Q = 3, 2

    from sonic_b200 import api

    good = sb.Assignment([1, 2, 3], [1, 2, 3], [1, 4, 9])
    api._check_prove_inputs(FakeHandle, good, list(range(1, 13)))
    for a, r in ((sb.Assignment([1, 2], [1, 2, 3], [1, 4, 9]), list(range(1, 13))), (good, list(range(1, 12)))):
        with pytest.raises(sb.SonicError):
            api._check_prove_inputs(FakeHandle, a, r)


def test_header_compiles_as_strict_c11_and_the_consumer_links():
    """tests/c/abi_multi.c names every prototype of include/sonic_b200.h with its full type and is built with
    -std=c11 -Wall -Wextra -Werror against the library: a declaration drifting from the definition (or from what a
    `foreign import ccall` would bind) stops the build.  Running it needs GPUs (tests/test_gpu_multi.py)."""
    import subprocess

    cdir = os.path.join(ROOT, "tests", "c")
    subprocess.run(["make", "-C", cdir, "-B", "abi_multi"], check=True, capture_output=True, text=True)
    assert os.path.exists(os.path.join(cdir, "abi_multi"))
    # the table in the C file covers the header: same symbol set
    text = open(os.path.join(cdir, "abi_multi.c")).read()
    for name in _declared_symbols():
        assert re.search(r"\b%s\b" % name, text), f"{name} is not referenced by tests/c/abi_multi.c"
    # without a GPU the program fails loudly with the library's own message
    import torch

    if not torch.cuda.is_available():
        out = subprocess.run([os.path.join(cdir, "abi_multi"), "1", "4"], capture_output=True, text=True)
        assert out.returncode == 2 and "no CPU fallback" in out.stderr


def test_header_documents_every_option_the_library_accepts():
    """sonic_set_option's names live in capi.cu; the header comment is the only documentation a binding author sees."""
    import re
    src = open(os.path.join(ROOT, "sonic_b200", "csrc", "capi.cu")).read()
    header = open(os.path.join(ROOT, "include", "sonic_b200.h")).read()
    body = src[src.index("int sonic_set_option("):]
    body = body[:body.index("\n}\n")]
    names = sorted(set(re.findall(r'!strcmp\(name, "([a-z_0-9]+)"\)', body)))
    assert len(names) >= 15, names
    missing = [n for n in names if '"%s"' % n not in header]
    assert not missing, "options accepted by sonic_set_option but absent from include/sonic_b200.h: %s" % missing
