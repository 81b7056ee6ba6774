"""World-size-2 coverage of the N>1 host logic on CPU (gloo): shard ranges, unique-id distribution, max-over-ranks
timing, whole-job rate, and the reference arm's "rank 0 only" rule.  The data-path collective itself is NCCL inside
libee_b200.so and is exercised on GPUs by bench.py --gpus N."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["EE_ROOT"])
import torch.distributed as dist
from ephemeris_explorer_b200 import distributed as eed
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
uid = bytes(range(128)) if rank == 0 else None
got = eed.broadcast_unique_id(dist, uid)
tg, sg = eed.shard_ranges(65536, world, rank, "allgather")
tr, sr = eed.shard_ranges(65536, world, rank, "allreduce")
mx = eed.max_over_ranks(dist, 10.0 + rank)
with open(os.path.join(os.environ["EE_OUT"], "rank%d.json" % rank), "w") as f:
    json.dump({"rank": rank, "uid_ok": got == bytes(range(128)), "tg": tg, "sg": sg, "tr": tr, "sr": sr, "mx": mx}, f)
dist.destroy_process_group()
'''


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_host_logic_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, EE_ROOT=str(ROOT), EE_NO_AUTOBUILD="1", EE_OUT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = [json.loads((tmp_path / ("rank%d.json" % r)).read_text()) for r in range(2)]
    assert len(rows) == 2
    assert all(r["uid_ok"] for r in rows)
    assert all(r["mx"] == 11.0 for r in rows)  # max over ranks, same on every rank
    # allgather: targets partition the bodies, every rank sees all sources
    assert [r["tg"] for r in rows] == [[0, 32768], [32768, 65536]] and all(r["sg"] == [0, 65536] for r in rows)
    # allreduce: sources partition the bodies, every rank owns all targets
    assert [r["sr"] for r in rows] == [[0, 32768], [32768, 65536]] and all(r["tr"] == [0, 65536] for r in rows)


def test_shard_ranges_edge_cases():
    from ephemeris_explorer_b200 import distributed as eed
    assert eed.shard_ranges(10, 1, 0, "allgather") == ((0, 10), (0, 10))
    with pytest.raises(ValueError):
        eed.shard_ranges(10, 3, 0, "allgather")
    with pytest.raises(ValueError):
        eed.shard_ranges(8, 2, 0, "ring")
    assert eed.whole_job_rate(65536, 8, 10, 2.0, sharded=True) == 65536 * 10 / 2.0
    assert eed.whole_job_rate(32, 8, 10, 2.0, sharded=False) == 32 * 10 * 8 / 2.0


def test_reference_arm_runs_on_rank0_only(tmp_path):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
