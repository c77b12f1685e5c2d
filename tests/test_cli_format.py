"""CPU: the drop-in binary's fast text formatter gives printf("%f") byte for byte (3.7 million values incl. exact
ties such as 15/128 and neighbours of every half-way point)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_format_unit_f_matches_printf(tmp_path):
    exe = str(tmp_path / "cli_format_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I",
                           os.path.join(ROOT, "ngsf-hmm_b200", "host", "cli"), "-o", exe,
                           os.path.join(ROOT, "tests", "cli_format_check.cpp")])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "0 mismatches" in p.stdout
