"""CPU-only checks: the C-ABI library loads, exports every symbol include/sapt.h declares, the
FFI struct mirrors match, Philox matches the Random123 known-answer vectors, and the host
logic (ladder) matches the golden fixtures. No compute call is made without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import golden_util as gu
from oracle import sapt_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_header_symbols():
    from sapt_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "sapt.h")).read()
    declared = set(re.findall(r"\b(sapt_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"sapt_problem_t"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"libsapt.so does not export {name}"
    assert set(_lib.EXPORTS) == declared
    assert L.sapt_abi_version() == 1


def test_error_convention_no_gpu_needed():
    from sapt_b200 import _lib
    L = _lib.lib()
    rc = L.sapt_prior(None, 0, 1, 1, 25.0, None, None)
    assert rc == _lib.ERR_INVALID
    assert b"sapt_prior" in L.sapt_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "sapt_prior")


PHILOX_KAT = [  # Random123 kat_vectors, philox4x32-10
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_known_answers(ctr, key, want):
    from sapt_b200 import _lib
    L = _lib.lib()
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    L.sapt_philox_host(c, k, o)
    assert tuple(o) == want
    got = so.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
    assert tuple(int(v) for v in got) == want


def test_ladder_matches_reference_fixture():
    from sapt_b200 import assign_temperatures, default_beta_ladder
    g = gu.load("swap_ladder")
    for R, mt in ((8, 2.0), (10, 4.0), (4, 4.0), (16, 6.0)):
        np.testing.assert_allclose(assign_temperatures(R, mt), g[f"ladder_R{R}_T{mt}"], rtol=1e-14)
    with pytest.raises(ValueError):
        default_beta_ladder(2, ntemps=4, Tmax=0.5)
    np.testing.assert_allclose(assign_temperatures(4, 4.0, geometric=False), [1, 2, 3, 4])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sapt_b200.problem import Problem
    with pytest.raises(RuntimeError):
        Problem([2, 2, 2], np.zeros((4, 3)))


def test_header_is_plain_c_and_c_host_builds(tmp_path):
    """include/sapt.h is the drop-in boundary: it must compile as C (C11) without any C++/CUDA/torch type, and the
    plain-C host example must link against libsapt.so."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include "sapt.h"\nint main(void) { sapt_step_args a; (void)a; return (int)sizeof(sapt_accept_args) == 0; }\n')
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], check=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples", "c_host"), "-B", "run_pt"], check=True, capture_output=True)
    assert os.path.isfile(os.path.join(ROOT, "examples", "c_host", "run_pt"))
