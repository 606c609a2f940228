"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol declared in
include/ofdm_b200.h, and refuses to create a context without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "ofdm_b200.h")).read()
    return sorted(set(re.findall(r"OFDM_API\s+[\w\s\*]+?\b(ofdm_\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from ofdm_b200 import _cabi
    return _cabi


def test_header_symbols_all_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 45
    so = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(so, n), f"{n} declared in the header but not exported"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(lib.SIGNATURES) == set(names)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    so = lib.load()
    h = ctypes.c_void_p()
    assert so.ofdm_ctx_create(ctypes.byref(h), 0, 0) == -3     # OFDM_ERR_NODEVICE
    import ofdm_b200
    with pytest.raises(ofdm_b200.OfdmError):
        ofdm_b200.Context(0)


def test_constellation_table_matches_oracle(lib):
    import numpy as np
    import ofdm_b200 as G
    import oracle as O
    for name in ("BPSK", "QPSK", "8PSK", "16QAM"):
        d, bps = G.constellation_func(name)
        dr, bpsr = O.constellation_func(name)
        assert bps == bpsr and np.max(np.abs(d - dr)) < 1e-15


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ofdm-course_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in re.sub(r'""".*?"""', "", open(os.path.join(pkg, fn)).read(), flags=re.S).replace("no CPU fallback", "")


def test_every_environment_switch_of_the_library_is_documented():
    """INTEGRATION.md section 8 lists the diagnostic switches: every getenv("OFDM_B200_...") in the CUDA sources must appear there
    (and nothing documented may have vanished from the sources)."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    used = set()
    for f in glob.glob(os.path.join(root, "ofdm-course_b200", "csrc", "*.cu")) + glob.glob(os.path.join(root, "ofdm-course_b200", "csrc", "*.cuh")):
        used |= set(re.findall(r'getenv\("(OFDM_B200_[A-Z0-9_]+)"\)', open(f).read()))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    section = doc[doc.index("## 8. Diagnostic switches"):]
    documented = set(re.findall(r"`(OFDM_B200_[A-Z0-9_]+)`", section))
    assert used == documented, (sorted(used - documented), sorted(documented - used))
