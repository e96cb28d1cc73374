"""Host I/O around the path and the kept CLI surface (run_real_data_script.py:18-62,82-93)."""
import gzip
import os
import struct

import numpy as np
import pytest

import run_real_data_script as cli
from multicomponent_t2_toolbox_b200 import nifti_io


def test_nifti_roundtrip_gz_and_plain(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.uniform(size=(5, 4, 3, 6))
    aff = np.array([[2.0, 0, 0, -10], [0, 2.5, 0, 5], [0, 0, 3.0, 1], [0, 0, 0, 1]])
    for name in ("x.nii.gz", "y.nii"):
        p = str(tmp_path / name)
        nifti_io.save(a, p, affine=aff)
        im = nifti_io.load(p)
        assert im.shape == a.shape
        assert np.array_equal(im.get_fdata(), a) and np.allclose(im.affine, aff)
    raw = gzip.open(str(tmp_path / "x.nii.gz")).read()
    assert struct.unpack("<i", raw[:4])[0] == 348 and raw[344:347] == b"n+1"
    assert struct.unpack("<h", raw[70:72])[0] == 64            # float64, like the reference's outputs
    assert len(raw) == 352 + a.size * 8
    assert raw[352:360] == struct.pack("<d", a[0, 0, 0, 0])      # Fortran order: x runs fastest
    assert raw[360:368] == struct.pack("<d", a[1, 0, 0, 0])


def test_nifti_int_mask_and_scaling(tmp_path):
    m = (np.arange(24).reshape(2, 3, 4) % 2).astype(np.int16)
    p = str(tmp_path / "m.nii.gz")
    nifti_io.save(m, p)
    im = nifti_io.load(p)
    assert np.array_equal(im.get_fdata(), m.astype(float))
    im.header["scl_slope"], im.header["scl_inter"] = 2.0, 1.0
    assert np.array_equal(im.get_fdata(), 2.0 * m + 1.0)
    with pytest.raises(ValueError):
        (tmp_path / "bad.nii").write_bytes(b"\x00" * 400)
        nifti_io.load(str(tmp_path / "bad.nii"))


def test_cli_flags_match_reference():
    argv = ("--path_to_folder /data/ --input Data.nii.gz --mask Mask.nii.gz --minTE 10.68 --nTE 32 --TR 1000 "
            "--FA_method spline --FA_smooth yes --denoise None --reg_method X2 --reg_matrix I --numcores -1 "
            "--myelin_T2=40 --savefig no --savefig_slice 30").split()   # --myelin_T2 relies on prefix matching
    a = cli.build_parser().parse_args(argv)
    assert a.myelin_T2_cutoff == 40.0 and a.reg_method == "X2" and a.FA_method == "spline" and a.nTE == 32
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(argv[:-2])                  # every flag is required, like the reference
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args([x if x != "X2" else "Tikhonov" for x in argv])
