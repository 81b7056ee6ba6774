"""Multi-GPU check as a pytest: runs scripts/multi_gpu_check.py under torchrun when at least two GPUs are visible
(sharded runs against the 1-GPU result: all-gather bitwise, all-reduce and the NVLink peer path <= 1e-12).  Skipped on
single-GPU boxes; bench.py --gpus N exercises the same paths there."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def test_sharded_runs_match_single_gpu():
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    g = 2 if g < 4 else (4 if g < 8 else 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(g), "--master-addr",
                          "127.0.0.1", "--master-port", str(port), str(ROOT / "scripts" / "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert '"ok": false' not in out.stdout
