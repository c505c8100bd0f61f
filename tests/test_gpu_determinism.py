"""Stand-in for a race check (compute-sanitizer is refused on this pool): the same batches through
different kernels, grids and streams-per-warp must give identical bytes.  One subprocess per set of
switches (they are read once per process): hash-chain kernel choice and resident CTAs for the
compressor, engine / lane-group size / table geometry for the decompressor."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {},
    {"BDF_HC_KERNEL": "old", "BDF_INFLATE_MODE": "group"},
    {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "0"},          # tables in shared memory (the default keeps them in global memory)
    {"BDF_LANE_CFG": "3"},
    {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "5"},          # global tables, width per block (the default only for large batches)
    {"BDF_HC_KERNEL": "new", "BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "1"},
    {"BDF_HC_KERNEL": "old", "BDF_HC_CTAS_PER_SM": "3", "BDF_INFLATE_MODE": "group", "BDF_INFLATE_GROUP": "32"},
    {"BDF_INFLATE_MODE": "auto", "BDF_INFLATE_SPLIT": "4", "BDF_LANE_WARPS": "3"},
    # without the first-block header pre-pass; the two engines side by side / lanes first
    {"BDF_INFLATE_PREHDR": "0", "BDF_INFLATE_SERIAL": "0"},
    {"BDF_INFLATE_MODE": "lane", "BDF_INFLATE_PREHDR": "0"},
    {"BDF_INFLATE_SERIAL": "2", "BDF_INFLATE_SPLIT": "2", "BDF_NOS_SPLIT": "0"},
    {"BDF_NOS_WAVE": "1"},
    # level 1: one match per round (the round-1 parse) on fewer resident CTAs; whole-window rounds without the bucket prefetch
    {"BDF_L1_WINDOW": "0", "BDF_L1_CTAS_PER_SM": "3"},
    {"BDF_L1_WINDOW": "1", "BDF_INFLATE_MODE": "group", "BDF_INFLATE_GROUP": "32"},
]


def run(env):
    e = dict(os.environ)
    for k in ("BDF_HC_KERNEL", "BDF_HC_CTAS_PER_SM", "BDF_INFLATE_MODE", "BDF_INFLATE_GROUP", "BDF_LANE_CFG",
              "BDF_INFLATE_SPLIT", "BDF_LANE_WARPS", "BDF_INFLATE_PREHDR", "BDF_INFLATE_SERIAL", "BDF_NOS_SPLIT", "BDF_NOS_WAVE",
              "BDF_L1_WINDOW", "BDF_L1_CTAS_PER_SM"):
        e.pop(k, None)
    e.update(env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "determinism_helper.py")], cwd=ROOT, env=e,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith(("compress", "decompress"))]


def test_same_bytes_whatever_the_kernel_choice():
    base = run(VARIANTS[0])
    assert len(base) == 13
    for v in VARIANTS[1:]:
        assert run(v) == base, v
