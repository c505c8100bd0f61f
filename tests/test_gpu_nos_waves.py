"""Levels 10..12 through several WAVES of the three-kernel pipeline (csrc/deflate_nos_split.cuh): with one
stream per SM and wave, 502 streams are four waves that reuse the same scratch slots; the bytes must
equal the single-kernel path's and a one-wave run's."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env):
    e = dict(os.environ)
    for k in ("BDF_NOS_SPLIT", "BDF_NOS_WAVE"):
        e.pop(k, None)
    e.update(env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "nos_wave_helper.py")], cwd=ROOT, env=e,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("digest")]


def test_waves_reuse_scratch_without_changing_bytes():
    one_wave = run({})
    assert len(one_wave) == 2
    assert run({"BDF_NOS_WAVE": "1"}) == one_wave
    assert run({"BDF_NOS_SPLIT": "0"}) == one_wave
