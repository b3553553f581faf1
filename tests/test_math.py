"""psd_math.h (the exp/log the kernels use) must be bit-identical to the libm the reference links
(glibc 2.39 x86-64 FMA variants).  Not GPU: the same header compiles for host and device."""
import os
import subprocess
import pytest
import oracle_bind
from helpers import ROOT


def test_exp_log_bit_identical_to_libm(tmp_path):
    if not oracle_bind.libm_matches_golden():
        pytest.skip("host libm is not glibc-2.39-FMA-identical; bit equality with it is not expected")
    exe = str(tmp_path / "math_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "math_check.c"), "-lm"])
    out = subprocess.run([exe, "1000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "0 mismatches; log:" in out.stdout and out.stdout.strip().endswith("0 mismatches")
