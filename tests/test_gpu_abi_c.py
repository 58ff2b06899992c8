"""GPU test of the boundary without Python in the way: tests/c/abi_smoke.c, a plain C99 program against include/h2agg.h,
is compiled with gcc, linked with libh2agg.so and run — init, upload, MSM, NTT, H fold, multi-open accumulation, KZG setup,
parameter file round trip, verifier view, transcript, each against known answers of the oracle's big-integer model."""
import os
import subprocess

import pytest

import halo2_aggregation_b200 as h2a

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_program_runs_the_hot_path(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(h2a.library_path())
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                        "-L" + libdir, "-lh2agg", "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, str(tmp_path / "smoke.params")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all ok" in r.stdout and "FAIL" not in r.stdout
