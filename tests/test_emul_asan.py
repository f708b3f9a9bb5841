"""CPU: AddressSanitizer + UBSan over the host builds of the phase-coded kernels (tests/emul/asan_main.cpp): tile
indices stay inside exactly the memory the CUDA launchers allocate as shared memory, for every latent dimension / K."""
import os, subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_phase_code_is_memory_clean_under_asan():
    out = os.path.join(ROOT, "tests", "emul", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "asan_main")
    cc = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-std=c++17", "-I",
                         os.path.join(ROOT, "svae_b200", "csrc"), os.path.join(ROOT, "tests", "emul", "asan_main.cpp"),
                         "-o", exe], capture_output=True, text=True)
    if cc.returncode != 0 and "sanitize" in (cc.stderr or "").lower() and "cannot find" in cc.stderr.lower():
        pytest.skip("libasan not installed")
    assert cc.returncode == 0, cc.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert run.returncode == 0 and "asan run ok" in run.stdout, (run.stdout + run.stderr)[-3000:]
