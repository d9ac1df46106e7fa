"""peaq_math.cuh (the kernels' ln / exp) on the host: same IEEE operation sequence as on the
device (only the reciprocal seed differs), checked against long double libm over 2^22 arguments
per function, plus the special values that must take the library path."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_log_exp_accuracy(tmp_path):
    exe = str(tmp_path / "fast_math_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-DPEAQ_MATH_HOST",
                           "-I", os.path.join(ROOT, "gstpeaq_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "fast_math_check.cpp")])
    out = subprocess.check_output([exe], text=True)
    exp_rel = float(re.search(r"exp max rel ([0-9.e+-]+)", out).group(1))
    log_rel = float(re.search(r"log max rel ([0-9.e+-]+)", out).group(1))
    assert exp_rel < 3e-16, out
    assert log_rel < 5e-16, out
    # result classes of the library functions are preserved
    assert "log(0)=-inf" in out and "log(inf)=inf" in out and "log(nan)=nan" in out
    assert re.search(r"log\(-1\)=-?nan", out)
    assert "exp(inf)=inf" in out and "exp(nan)=nan" in out and "exp(-1e+06)=0" in out or "exp(-1e6)=0" in out
    assert "exp(-745)=4.94066e-324" in out     # denormal results through the library path
    assert "log(1e-310)=-713.801" in out       # denormal arguments too
