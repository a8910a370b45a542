"""CPU: emfusion_b200/csrc/seq_add.h (n repeated fp32 additions in O(1), used by the raycast jumps) must equal the
sequential loop bit for bit -- metric ray lengths, rounding ties, zero crossings, random bit patterns."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_seq_add_matches_sequential_loop(tmp_path):
    exe = tmp_path / "seq_add_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "emfusion_b200", "csrc"),
                    os.path.join(ROOT, "tests", "csrc", "seq_add_check.c"), "-o", str(exe), "-lm"], check=True)
    r = subprocess.run([str(exe), "400000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert " bad 0 " in r.stdout
