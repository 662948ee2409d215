"""CPU tests of the C-ABI boundary: the shared library loads without a GPU, exports every symbol that include/bgarena.h
declares (and nothing undeclared with the bg_ prefix), and fails loudly -- never falls back -- when no CUDA device exists."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "bgarena.h")
SO = os.path.join(ROOT, "mlp-ppo-2ply-multi_b200", "libbgarena.so")


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def so():
    if not os.path.exists(SO):
        subprocess.run(["bash", os.path.join(ROOT, "mlp-ppo-2ply-multi_b200", "csrc", "build.sh")], check=True)
    return C.CDLL(SO)


def test_exports_every_declared_symbol(so):
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(so, n), f"{n} declared in include/bgarena.h but not exported"


def test_no_undeclared_bg_exports():
    out = subprocess.run(["nm", "-D", "--defined-only", SO], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\b(bg_[a-z0-9_]+)\b", out)))
    assert exported == declared_symbols()


def test_python_binding_matches_header():
    import mlp_ppo_2ply_multi_b200 as bg

    assert sorted(bg._lib.SIGNATURES) == declared_symbols()
    assert bg._lib.lib().bg_abi_version() == 1


def test_fails_loudly_without_gpu(so):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    so.bg_device_count.restype = C.c_int32
    assert so.bg_device_count() == 0
    so.bg_encode.restype = C.c_int32
    so.bg_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    buf = (C.c_char * 4096)()
    rc = so.bg_encode(C.addressof(buf), C.addressof(buf), 1, C.addressof(buf), None)
    assert rc == -2  # BG_ERR_CUDA: no CPU fallback
    so.bg_last_error.restype = C.c_char_p
    assert b"no CUDA device" in so.bg_last_error()
    import mlp_ppo_2ply_multi_b200 as bg

    with pytest.raises(ValueError):
        bg.encode(torch.zeros((1, 52), dtype=torch.int8), torch.zeros(1, dtype=torch.uint8))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU: one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-positions", "1024"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "afterstates_evaluated_per_sec" and line["unit"] == "afterstates/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "afterstates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]
