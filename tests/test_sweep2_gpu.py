"""The two-rows-per-lane triangular sweep (csrc/ifl_sweep2.cuh, opt-in with IFL_SWEEP2=1) against the oracle.

The kernel choice is read once per process, so the F3 parity tests of test_parity_gpu.py (applyPreconditioner F3:495 in
isolation, whole PCG timesteps bit for bit in strict mode on square and ragged grids, the 100-step fast-path run) are
re-run in a child pytest process with the variable set."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_row_sweep_passes_the_f3_parity_tests():
    env = dict(os.environ, IFL_SWEEP2="1")
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_parity_gpu.py"), "-x", "-q", "-m", "gpu",
           "-k", "pcg or fast_path or precon or stage or projection", "-p", "no:cacheprovider"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
