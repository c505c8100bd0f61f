"""The -DBDF_CHECK build (device-side bounds / invariant assertions at the kernels' memory indices)
over good, corrupted, truncated and short-room batches: no assertion may fire.  One subprocess per
engine setting (the library is chosen when the package is imported)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{}, {"BDF_INFLATE_MODE": "lane"}, {"BDF_INFLATE_MODE": "group", "BDF_HC_KERNEL": "new"}],
                         ids=["default", "lane", "group-hcs"])
def test_no_device_assertion_fires(env):
    from libdeflate_rsx_b200 import build
    lib = build.build(check=True)
    e = dict(os.environ)
    e.update(env)
    e["BDF_LIBRARY"] = lib
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "check_build_helper.py")], cwd=ROOT, env=e,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "check word 0" in out.stdout, out.stdout[-500:]
