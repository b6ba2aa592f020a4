"""SURVEY.md section 8f rows on the CPU: FITS edge (pure host code) and the accretor / WVT helpers on
the SIMT-emulator engine (see tests/test_emu_kernels.py for what that is)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import EmuEngine  # noqa: E402
import frow_checks  # noqa: E402


def test_fits_round_trip(tmp_path):
    from tess_b200.fitsio import read_image, write_image
    rng = np.random.default_rng(0)
    for dt in ("uint8", "int16", "int32", "int64", "float32", "float64"):
        a = (rng.uniform(0, 100, (7, 13)) * (1 if dt.startswith("f") else 1)).astype(dt)
        p = str(tmp_path / ("a_%s.fits" % dt))
        write_image(p, a, header={"OBJECT": "tess test", "TARGSN": 50.0, "NBINS": 271, "FLAG": True})
        assert os.path.getsize(p) % 2880 == 0
        b, h = read_image(p)
        assert b.dtype == np.dtype(dt) and b.dtype.isnative and np.array_equal(a, b)
        assert h["NAXIS1"] == 13 and h["NAXIS2"] == 7 and h["OBJECT"] == "tess test" and h["TARGSN"] == 50.0
        assert h["NBINS"] == 271 and h["FLAG"] is True
    # bool / odd ints are widened; 3-D cube; BZERO/BSCALE applied on read
    p = str(tmp_path / "c.fits")
    write_image(p, np.arange(24, dtype=np.uint16).reshape(2, 3, 4), header={"BZERO": 32768.0, "BSCALE": 1.0})
    b, h = read_image(p)
    assert b.shape == (2, 3, 4) and b.dtype == np.float64 and b[1, 2, 3] == 23 + 32768.0
    raw, _ = read_image(p, raw=True)
    assert raw.dtype == np.int64 and raw[1, 2, 3] == 23
    with open(str(tmp_path / "bad.fits"), "wb") as f:
        f.write(b"not fits" * 400)
    with pytest.raises(ValueError):
        read_image(str(tmp_path / "bad.fits"))


@pytest.mark.skipif(not os.path.exists("/root/reference/scripts/iso_sn.fits"), reason="reference tree not mounted")
def test_fits_reads_the_reference_segmap():
    """scripts/iso_sn.fits: the accretor segmentation map written by astropy in
    tess/tests/test_pixel_accretion.py:35 (64x64 int64, SURVEY.md section 8 note on C1)."""
    from tess_b200.fitsio import read_image
    seg, h = read_image("/root/reference/scripts/iso_sn.fits")
    assert seg.shape == (64, 64) and seg.dtype == np.int64 and h["BITPIX"] == 64
    assert seg.min() >= -1 and seg.max() < 64 * 64


def test_accretion_helpers():
    frow_checks.check_accretion_helpers(EmuEngine)


def test_accretion_helpers_against_reference_fixture():
    frow_checks.check_accretor_fixture(EmuEngine)


def test_wvt_binning_to_target_sn():
    frow_checks.check_wvt_binning(EmuEngine, N=56, k=10)
