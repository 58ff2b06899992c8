"""GPU test of the host side above the C ABI in C++: examples/simple_example.cpp restates the reference's
examples/simple-example.rs with include/h2agg.hpp — Setup::new from the XorShift seed, keygen from the copy constraints of
MyCircuit at k = 9, create_proof, verify_proof -> [e, f, w, zw], commit_lagrange over verifier_params, the 40 public inputs of the
aggregation circuit, a parameter file round trip and the mul_var witness cells of s * W — and checks every step against the
oracle's known answers (examples/simple_example_kat.h): the proof bytes are the oracle prover's, byte for byte."""
import os
import subprocess

import pytest

import halo2_aggregation_b200 as h2a

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_restatement_of_the_reference_example(tmp_path):
    exe = str(tmp_path / "simple_example")
    libdir = os.path.dirname(h2a.library_path())
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "simple_example.cpp"), "-L" + libdir, "-lh2agg", "-Wl,-rpath," + libdir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, str(tmp_path / "agg.params")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all ok" in r.stdout and "FAIL" not in r.stdout and "proof size is 1248" in r.stdout, r.stdout
