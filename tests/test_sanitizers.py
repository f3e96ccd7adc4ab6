"""Memory-safety check of the engine source: the device code compiled at warp width 1 with ASan + UBSan runs every
model family (tests/hostemu/asan_run.py).  compute-sanitizer is closed on the GPU pool, so this is the tool that
watches table bounds, arena bookkeeping and record logs; the CUDA build is the same source at warp width 32."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def runtime(name):
    path = subprocess.run(["gcc", f"-print-file-name={name}"], capture_output=True, text=True).stdout.strip()
    return path if os.path.isabs(path) and os.path.exists(path) else None


@pytest.mark.parametrize("defines", [[], ["-DITSB_HIDDEN_FAST=1"]], ids=["as k_run", "as k_run_in_place"])
def test_every_model_family_is_clean_under_asan_and_ubsan(tmp_path, defines):
    asan, ubsan = runtime("libasan.so"), runtime("libubsan.so")
    if not asan or not ubsan:
        pytest.skip("gcc sanitizer runtimes not installed")
    lib = str(tmp_path / "libitsb_hostemu_asan.so")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                    "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
                    *defines, "-o", lib, os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")], check=True)
    env = dict(os.environ, LD_PRELOAD=f"{asan}:{ubsan}", ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "hostemu", "asan_run.py"), lib], env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.strip().endswith("clean")
    assert "runtime error" not in out.stderr and "AddressSanitizer" not in out.stderr
