"""CPU: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "scint_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    from scintools_b200 import _lib
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(_lib.lib, n), "missing export %s" % n
    assert set(_lib.EXPORTS) == set(names)
    assert _lib.lib.sb_abi_version() >= 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scintools_b200 import _device
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _device.device()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "scintools_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn


def test_host_axes_match_oracle():
    import numpy as np
    from oracle import thth_oracle as TO
    from scintools_b200 import ththmod as thth
    t = np.arange(150) * 28.7
    f = 1400 + 0.0321 * np.arange(64)
    for pad in (0, 3):
        assert np.array_equal(thth.fft_axis(t, "mHz", pad), TO.fft_axis(t, "mHz", pad))
        assert np.array_equal(thth.fft_axis(f, "us", pad), TO.fft_axis(f, "us", pad))
    e = np.linspace(-0.4, 0.4, 512)
    assert np.array_equal(thth.theta_centres(e), TO.theta_centres(e))
    fd, tau = TO.fft_axis(t, "mHz"), TO.fft_axis(f, "us")
    assert np.array_equal(thth.min_edges(0.3, fd, tau, 50.0), TO.min_edges(0.3, fd, tau, 50.0))
