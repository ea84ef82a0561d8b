"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/orbx.h declares; host-side tables
match the oracle; compute calls fail loudly (no CPU fallback) when no CUDA device is visible."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as graft
from oracle import binding as ob
from orb_slam2_commit_b200 import ORBextractor, OrbxError, api, hamming_top2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    graft.build()


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "orbx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    L = C.CDLL(api.library_path())
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert L.orbx_abi_version() == 1


def test_keypoint_layout_is_cv_keypoint():
    assert api.KP_DTYPE.itemsize == 28
    assert [api.KP_DTYPE.fields[f][1] for f in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


@pytest.mark.parametrize("cfg", [(1000, 1.2, 8), (2000, 1.2, 8), (1200, 1.2, 8), (8000, 1.2, 12), (500, 1.5, 4), (300, 1.1, 6)])
def test_ctor_tables_match_oracle(cfg):
    nf, sc, nl = cfg
    e = ORBextractor(nf, sc, nl, 20, 7)
    o = ob.Extractor(nf, sc, nl, 20, 7).tables()
    assert e.GetLevels() == nl and e.GetScaleFactor() == np.float32(sc)
    for mine, key in ((e.GetScaleFactors(), "scale_factors"), (e.GetInverseScaleFactors(), "inv_scale_factors"),
                      (e.GetScaleSigmaSquares(), "sigma2"), (e.GetInverseScaleSigmaSquares(), "inv_sigma2")):
        assert np.array_equal(mine.view(np.uint32), o[key].view(np.uint32)), key
    assert np.array_equal(e.features_per_level(), o["features_per_level"])
    assert np.array_equal(e.umax(), o["umax"])


def test_invalid_arguments():
    with pytest.raises(OrbxError):
        ORBextractor(100, 1.2, 0, 20, 7)
    with pytest.raises(OrbxError):
        ORBextractor(100, 1.2, 17, 20, 7)
    with pytest.raises(OrbxError):
        ORBextractor(100, 1.0, 8, 20, 7)


def test_empty_image_is_silent():
    e = ORBextractor(100, 1.2, 2, 20, 7)
    kps, desc = e(np.zeros((0, 0), np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
    n = C.c_int32(-5)
    assert api.lib().orbx_extract(e._h, None, 0, 0, 0, None, 0, C.byref(n), None) == 0 and n.value == 0


def test_no_cpu_fallback_without_device():
    if api.lib().orbx_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    e = ORBextractor(100, 1.2, 2, 20, 7)
    with pytest.raises(OrbxError) as ei:
        e(np.zeros((240, 320), np.uint8))
    assert ei.value.code == 4
    with pytest.raises(OrbxError):
        hamming_top2(np.zeros((4, 32), np.uint8), np.zeros((8, 32), np.uint8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "orb_slam2_commit_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "orb_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_geometry_the_reference_cannot_run_is_rejected():
    """Portrait levels (nIni == 0, ORBextractor.cc:567), levels under 62 px (nCols == 0, :846), > 4096 px and
    > 2040 features on one level return ORBX_ERR_UNSUPPORTED from pure host arithmetic (no device needed)."""
    L = api.lib()
    cases = [((500, 1.2, 2), 240, 640), ((500, 1.2, 8), 120, 100), ((500, 1.2, 2), 5000, 3000), ((60000, 1.2, 8), 640, 480)]
    for (nf, sc, nl), w, h in cases:
        e = ORBextractor(nf, sc, nl, 20, 7)
        rc = L.orbx_reserve(e._h, w, h, 1)
        assert rc == 2, (w, h, rc, L.orbx_last_error())
    e = ORBextractor(500, 1.2, 8, 20, 7)
    assert L.orbx_reserve(e._h, 0, 480, 1) == 1 and L.orbx_reserve(e._h, 640, 480, 0) == 1
