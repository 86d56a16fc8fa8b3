"""The C ABI (include/ebcadrl.h): header <-> ctypes <-> libebcadrl.so consistency.  CPU only,
no compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

import oracle_backend as ob
from ebc import abi
from ebc.config import SimConfig

ROOT = ob.ROOT
HEADER = os.path.join(ROOT, "include", "ebcadrl.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(abi.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return abi.load()


def declared(prefix_ref):
    src = open(HEADER).read()
    names = set(re.findall(r"\b(ebc_(?:ref_)?[a-z_0-9]+)\s*\(", src))
    return {n for n in names if n.startswith("ebc_ref_") == prefix_ref}


def test_library_exports_every_declared_symbol(lib):
    want = declared(False)
    assert want == set(abi.PROTOTYPES), want ^ set(abi.PROTOTYPES)
    for name in want:
        assert hasattr(lib, name), name
    assert lib.ebc_abi_version() == abi.ABI_VERSION


def test_oracle_exports_every_ref_symbol(oracle):
    for name in declared(True):
        assert hasattr(oracle.lib, name), name


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu\\n",'
                   'sizeof(ebc_config),sizeof(ebc_state),sizeof(ebc_weights),offsetof(ebc_config,time_step),'
                   'offsetof(ebc_config,orca_neighbor_dist),offsetof(ebc_weights,mlp3));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(abi.EbcConfig), ctypes.sizeof(abi.EbcState), ctypes.sizeof(abi.EbcWeights),
            abi.EbcConfig.time_step.offset, abi.EbcConfig.orca_neighbor_dist.offset, abi.EbcWeights.mlp3.offset]
    assert got == want


def test_no_cpu_fallback(lib):
    """Without a GPU the product must fail loudly, never compute on the host."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = SimConfig().to_abi(4, 3, 0, 0, 81)
    h = abi.SIM()
    rc = lib.ebc_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc == -3 and b"no CUDA device" in lib.ebc_last_error(None)
    from ebc.engine import BatchedSim
    with pytest.raises(abi.EbcError):
        BatchedSim(SimConfig(), 4, 3, device="cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "eb-cadrl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "ebc_ref_" not in text and "libebc_oracle" not in text and "oracle_backend" not in text, f
