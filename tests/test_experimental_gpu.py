"""Variants that are compiled in but OFF by default because they have not run on a GPU yet (see notes/README.md).
Skipped unless B2C_RUN_EXPERIMENTAL=1; each runs the ordinary parity cases in a subprocess with the variant's switch on."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(os.environ.get("B2C_RUN_EXPERIMENTAL") != "1", reason="experimental variants run only on request")
@pytest.mark.parametrize("switch", ["B2C_WGRAD_COMPACT", "B2C_WGRAD3_TMA"])
def test_variant_passes_the_parity_cases(switch):
    env = dict(os.environ, **{switch: "1"})
    env.pop("B2C_RUN_EXPERIMENTAL")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                        "-k", "s2 or 1x1 or 3x3"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
