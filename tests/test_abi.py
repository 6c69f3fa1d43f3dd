"""The C-ABI library loads without a GPU and exports every symbol include/mfb200.h declares, with the signature
table the ctypes binding uses; compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from mirrorfusion_b200 import _lib
    return _lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mfb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mfb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mfb200.h but not exported"
    assert sorted(built.EXPORTS) == declared, "ctypes signature table and header disagree"


def test_abi_version_and_error_channel(built):
    lib = built.load()
    assert lib.mfb_abi_version() == 1
    import torch
    if not torch.cuda.is_available():
        rc = lib.mfb_init(0)
        assert rc != 0
        msg = lib.mfb_last_error().decode()
        assert "no CUDA device" in msg or "failed" in msg


def test_struct_layout_matches_header(built):
    # mfb_conv_desc: 7 ints, ptr, int, 3 ptrs, 3 ints, ptr x3, int, ptr x4, 2 ints — compare against a C compile of the header
    import subprocess, tempfile, textwrap
    code = textwrap.dedent("""
        #include <stdio.h>
        #include <stddef.h>
        #include "mfb200.h"
        int main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(mfb_conv_desc), offsetof(mfb_conv_desc, x),
                                offsetof(mfb_conv_desc, w), offsetof(mfb_conv_desc, out), offsetof(mfb_conv_desc, igemm_mode)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(code)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    D = built.ConvDesc
    assert got == [ctypes.sizeof(D), D.x.offset, D.w.offset, D.out.offset, D.igemm_mode.offset]


def test_product_path_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mirrorfusion_b200 import ops
    with pytest.raises(built.MfbError):
        ops.lib()
    from mirrorfusion_b200.schedulers import B200UniPCScheduler
    s = B200UniPCScheduler()
    s.set_timesteps(4)
    with pytest.raises(RuntimeError):
        s.step(torch.zeros(1, 4, 8, 8), s.timesteps[0], torch.zeros(1, 4, 8, 8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "reflecting-reality_b200", "mirrorfusion_b200")
    import re
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"
