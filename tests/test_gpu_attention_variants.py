"""-m gpu, EXPERIMENTAL (skipped unless FD_TEST_EXPERIMENTAL=1): opt-in variants of the long-KV attention kernel, each in its own
process because the variant is chosen from the environment once per process.  FD_ATTN_POLY=3 (part of the exponentials as a
polynomial on the FMA pipe) has been run on B200 (correct, +2 %); FD_ATTN_W8=1 (eight softmax warps) has NOT been run yet —
run it under `timeout`: a protocol error in it would be a hang, not a wrong number."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FD_TEST_EXPERIMENTAL") != "1", reason="opt-in path; set FD_TEST_EXPERIMENTAL=1")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"FD_ATTN_POLY": "3"}, {"FD_ATTN_W8": "1"}, {"FD_ATTN_W8": "1", "FD_ATTN_POLY": "3"}])
def test_attention_variant_matches_sdpa(env):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dev_attn_poly.py")], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    errs = [float(m) for m in re.findall(r"rel-L2 vs SDPA fp32 ([0-9.e+-]+)", r.stdout)]
    assert len(errs) == 4 and max(errs) < 5e-3, r.stdout
