"""The restated glibc log (devastator_b200/csrc/deva_log.cuh, host build) against the host libm's
std::log, bit for bit - the arithmetic the reference's PHOLD delay depends on (test/phold.cxx:114)."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_log_restatement_is_bit_exact(tmp_path):
    exe = str(tmp_path / "log_sweep")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-pthread",
                           os.path.join(ROOT, "tests", "csrc", "log_sweep.cpp"), "-o", exe])
    out = json.loads(subprocess.check_output([exe, "3000000", str(min(8, os.cpu_count() or 1))]))
    assert out["log_mismatch"] == 0 and out["dt_mismatch"] == 0 and out["edge_mismatch"] == 0, out


def test_log_table_matches_host_libm(tmp_path):
    """The committed table is what tools/extract_glibc_log_table.py reads out of this host's libm."""
    out = str(tmp_path / "tab.inc")
    subprocess.check_call(["python", os.path.join(ROOT, "tools", "extract_glibc_log_table.py"),
                           "/lib/x86_64-linux-gnu/libm.so.6", out], stdout=subprocess.DEVNULL)
    strip = lambda p: [l for l in open(p).read().splitlines() if not l.startswith("//")]
    assert strip(out) == strip(os.path.join(ROOT, "devastator_b200", "csrc", "glibc_log_data.inc"))
