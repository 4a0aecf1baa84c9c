"""-m gpu: the opt-in variant of the long-KV attention kernel (own process: the variant is chosen from the environment once per
process).  FD_ATTN_POLY=3 = 3 of every 8 exponential pairs as a polynomial on the FMA pipe: correct, +2 % at 4096 keys, nothing at
1024 (profiles/r02_attention_variants.txt).  The eight-softmax-warp variant of round 1 lost its A/B (712 vs 702 TFLOP/s) and was
removed."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"FD_ATTN_POLY": "3"}])
def test_attention_variant_matches_sdpa(env):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dev_attn_poly.py")], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    errs = [float(m) for m in re.findall(r"rel-L2 vs SDPA fp32 ([0-9.e+-]+)", r.stdout)]
    assert len(errs) == 4 and max(errs) < 5e-3, r.stdout
