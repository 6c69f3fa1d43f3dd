"""The C ABI from a host with no Python and no torch: tests/c_host/abi_demo.c (plain C, include/mfb200.h + libmfb200.so + cudart)
builds a linear + LayerNorm as a recorded program (mfb_program_begin / _end), replays it with ONE call (mfb_program_run) and checks
it against a scalar C evaluation."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "reflecting-reality_b200", "mirrorfusion_b200", "lib")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_demo(tmp_path):
    exe = str(tmp_path / "abi_demo")
    subprocess.run(["gcc", "-std=c11", "-O1", os.path.join(ROOT, "tests", "c_host", "abi_demo.c"), "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(CUDA, "include"), "-L", LIBDIR, "-lmfb200", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm",
                    f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}", "-o", exe], check=True)
    return exe


@pytest.mark.gpu
def test_c_host_runs_a_recorded_program(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    r = subprocess.run([build_demo(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C_HOST_OK" in r.stdout, (r.stdout, r.stderr)


def test_c_host_links_and_fails_loudly_without_a_gpu(tmp_path):
    """`-m "not gpu"`: the demo compiles as C against the header, links against the in-tree library, and — where no GPU is visible —
    stops at mfb_init with the library's error text instead of computing anything on the CPU."""
    import torch
    import __graft_entry__ as ge
    ge.build()
    exe = build_demo(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_c_host_runs_a_recorded_program")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "no CUDA device" in r.stderr and "C_HOST_OK" not in r.stdout
