"""
The four-step tapered-FFT kernel (csrc/mtm_4s.cu) is opt-in: the library reads SPYB_MTM_4S once per process, so its
parity cases (tests/cases_mtm_4s.py) run in a child interpreter with the switch set.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_four_step_kernel_cases():
    env = dict(os.environ, SPYB_MTM_4S="1")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "cases_mtm_4s.py"), "-q", "-x", "-m", "gpu",
                          "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
