"""The arithmetic behind csrc/solver_kernels.cu: div_tiny, restated in C and checked against IEEE division on the CPU
(no GPU needed).  Pockets of free pixels enclosed by depth-0 scribbles decay into a denormal limit cycle; their weighted
means (ref: src/GPUSolver.cu:104 `sum / count`, IEEE div.rn there) have numerators below 2^-100, which the sweep kernels
divide exactly through a scaled quotient plus a midpoint-tie correction.  This test pins both halves of that claim: with
the correction every quotient equals a / b bit for bit, without it about 1 % of them do not."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "div_tiny_check.c")
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")


def _run(count, fix_ties):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "div_tiny_check")
    r = subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", SRC, "-o", exe, "-lm"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([exe, str(count), str(fix_ties)], stdout=subprocess.PIPE, text=True, timeout=280).stdout.split()
    return int(out[0]), int(out[1])


def test_scaled_quotient_with_tie_correction_is_ieee_division():
    n, bad = _run(30_000_000, 1)
    assert n == 30_000_000 and bad == 0


def test_the_tie_correction_is_necessary():
    n, bad = _run(5_000_000, 0)
    assert bad > n // 1000            # double rounding through the denormal grid goes wrong on midpoint ties
