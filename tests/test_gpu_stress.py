"""One large randomised differential batch through the LARGE-batch decompress configuration (header
pre-pass + lane kernel with its tables in global memory and the litlen width chosen per block + lane
groups for the heavy streams): 34000 small streams per call, several zlib strategies, multi-block
streams, random capacities and corruptions; status and bytes equal the oracle's stream by stream.
(The small batches of test_gpu_fuzz.py run the small-batch configuration.)"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fmt", ["1", "2"])
def test_large_random_batch_matches_oracle(fmt):
    env = dict(os.environ, STRESS_FORMATS=fmt)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "gpurun_scripts", "stress_inflate.py"), "34000", "9"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "STRESS OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    assert "mismatches 0" in out.stdout
