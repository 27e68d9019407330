"""Runs every primitive check of tests/test_hostsim.py against the SAME device functions executed on a real B200
(tests/hostsim/libgpusim.so: one-thread kernels around each device function).  Localises a GPU-only failure
(PTX carry chains, stack, constant memory) to a single primitive."""
import ctypes
import os
import sys
import subprocess

import pytest

import test_hostsim as T

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "hostsim", "libgpusim.so")
SRC = os.path.join(HERE, "hostsim", "gpusim.cu")


@pytest.fixture(scope="module")
def hs():
    csrc = os.path.join(HERE, "..", "milagro_bls_b200", "csrc")
    gen = os.path.join(HERE, "hostsim", "make_gpusim.py")
    deps = [os.path.join(HERE, "hostsim", f) for f in ("hostsim.cpp", "test_only.cuh", "make_gpusim.py")]
    deps += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call([sys.executable, gen])          # regenerate gpusim.cu from hostsim.cpp
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                               "-shared", "-o", LIB, SRC])
    lib = ctypes.CDLL(LIB)
    lib.hs_g1_op.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p]
    lib.hs_g2_op.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p]
    return lib


test_fp_ops = T.test_fp_ops
test_fp_mont_mul_raw_edges = T.test_fp_mont_mul_raw_edges
test_fp_sqr_raw_limb_patterns = T.test_fp_sqr_raw_limb_patterns
test_fp_pow_sliding_window = T.test_fp_pow_sliding_window
test_fp2_ops = T.test_fp2_ops
test_fp2_sqrt_or_z = T.test_fp2_sqrt_or_z
test_fp12_ops = T.test_fp12_ops
test_hash_to_field_and_map = T.test_hash_to_field_and_map
test_sswu_and_iso3 = T.test_sswu_and_iso3
test_hash_to_g2 = T.test_hash_to_g2
test_g1_ops = T.test_g1_ops
test_g2_ops = T.test_g2_ops
test_pairing_gt_bytes_and_verify = T.test_pairing_gt_bytes_and_verify
