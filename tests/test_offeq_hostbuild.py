"""CPU check of the off-equatorial jet physics written for the rest of SURVEY 8 f4 (kgeo_b200/csrc/kg_offeq.cuh: metric, BZ monopole /
power-law field with its field-line speed, drift frame, redshift, tetrad, Penrose-Walker transport, per-crossing Stokes parameters).
The header's functions are __host__ __device__: it is compiled for the host with g++ and compared with the pinned oracle on the
crossings of tests/golden/jet_pipeline.npz (and with the unmodified reference's raw Stokes parameters stored there).  The GPU side of
the same comparison is tests/test_gpu_jet.py."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_offeq_device_functions_on_the_host(tmp_path):
    so = str(tmp_path / "_hostsim_offeq.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "kgeo_b200", "csrc"),
                    os.path.join(ROOT, "tests", "hostsim", "hostsim_offeq.cpp"), "-o", so], check=True, capture_output=True, timeout=300)
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
    import check_offeq
    worst = check_offeq.run(so, verbose=False)
    assert worst < 1e-10, worst
