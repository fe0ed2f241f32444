"""CPU: csrc/libm_exact.h (the operation-by-operation restatement of glibc's sinf / cosf that the device-resident OuterBnB uses to
build rotation matrices, jly_goicp.cpp:729-747) returns bit for bit what this host's C library returns.  The header is plain
C++ on the host, so the check needs no GPU: every 13th float in [0, 8] and its negative (6.7e8 comparisons were run exhaustively
once; rotation angles are <= sqrt(3)*pi = 5.45)."""
import os
import subprocess

from conftest import ROOT


def test_libm_exact_matches_host_libm(tmp_path):
    exe = str(tmp_path / "lme_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "go-icp-protein-cavities_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "lme_check.cpp"), "-lm"])
    # the variant glibc's ifunc rule selects on this host, then both variants forced: on this sample the FMA and the SSE2 build of
    # the C library round to the same floats, so each restated variant must agree with whichever one the host runs
    for extra in ([], ["0"], ["1"]):
        out = subprocess.run([exe, "13"] + extra, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:]
        last = out.stdout.strip().splitlines()[-1].split()
        assert last[0] == "fma_variant" and int(last[3]) > 3e8 and int(last[5]) == 0, out.stdout[-500:]
