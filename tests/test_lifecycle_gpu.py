"""Engine lifetime: b200_finalize releases every stream, event, staging buffer and scheduler counter, and the next call
re-initialises lazily with the same results (the reference's bli_init/bli_finalize pair may be cycled by an application,
frame/base/bli_init.c:87-99).  Runs in its own process so that the session-wide engine is left alone."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_finalize_then_reinit():
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "lifecycle_driver.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["before"] < 1e-12 and out["host_before"] < 1e-12, out
    for cycle in range(3):
        assert out[f"same_bits_{cycle}"], out
        assert out[f"host_{cycle}"] < 1e-12 and out[f"batch_{cycle}"] < 1e-12, out
    assert out["launches"] >= 4 * 2 + 3 * 3


def test_environment_presets_follow_the_blis_env_convention():
    """BLIS_B200_<KEY> presets a b200_set_option knob at initialisation (the reference reads BLIS_* variables once at init,
    frame/base/bli_env.c:68): BLIS_B200_DGEMM_CFG=6 must route an aligned dgemm to the cp.async kernel instead of the TMA one."""
    import os
    import subprocess
    import sys
    code = ("import torch; from blis_b200 import api; "
            "a=torch.rand(512,512,dtype=torch.float64,device='cuda'); c=torch.zeros_like(a); "
            "api.bli_dgemm(0,0,512,512,512,1.0,a,1,512,a,1,512,0.0,c,1,512); torch.cuda.synchronize(); print(api.last_kernel())")
    root = str(__import__("pathlib").Path(__file__).resolve().parent.parent)
    for env_extra, want in (({"BLIS_B200_DGEMM_CFG": "6"}, "gemm_dmma_ws_kernel<double,128x128x16"), ({"BLIS_B200_DGEMM_CFG": "9"}, "gemm_dmma_tma_kernel"),
                            ({"BLIS_B200_DGEMM_CFG": "nonsense"}, "gemm_dmma_")):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root, env=dict(os.environ, **env_extra))
        assert r.returncode == 0, r.stderr[-2000:]
        assert r.stdout.strip().splitlines()[-1].startswith(want), (env_extra, r.stdout)
