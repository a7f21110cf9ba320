"""fast_exp (ndtpso_slam_b200/csrc/fast_exp.h), compiled for the host: at most 1 ulp from glibc exp over
4 M arguments in the ranges the NDT score produces, exact specials (0, -inf, flush below -708, overflow, NaN)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_exp_within_one_ulp(tmp_path):
    exe = str(tmp_path / "fexp")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "native", "fast_exp_check.cpp"), "-lm"], check=True)
    res = subprocess.run([exe, "4000000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "max_ulp_err" in res.stdout
    assert float(res.stdout.split()[1]) <= 1.0


def test_exp_table_is_correctly_rounded():
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    vals = []
    for line in open(os.path.join(ROOT, "ndtpso_slam_b200", "csrc", "exp_table.inc")):
        line = line.strip()
        if line.startswith("0x"):
            vals.append(float.fromhex(line.split(",")[0]))
    assert len(vals) == 16
    for j, v in enumerate(vals):
        assert v == float(Decimal(2) ** (Decimal(j) / Decimal(16)))
