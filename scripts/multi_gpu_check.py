#!/usr/bin/env python3
"""Run under torchrun on G >= 2 GPUs: sharded n-body stepping against the same run on one GPU.

  allgather (targets sharded): bitwise identical to the single-GPU result, in throughput AND parity mode
  allreduce (sources sharded, the north-star exchange): <= 1e-12 relative (summation order differs)
Prints one JSON line per check on rank 0 and exits non-zero on any failure.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    import ephemeris_explorer_b200 as ee
    from ephemeris_explorer_b200 import distributed as eed
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    h = 2.0 ** -10
    for n, steps, modes in ((4096, 30, ((ee.MODE_THROUGHPUT, "throughput"), (ee.MODE_PARITY, "parity"))),
                            (32768, 14, ((ee.MODE_THROUGHPUT, "throughput-sym"),))):  # n >= 32768: pair-symmetric kernel
      pos, vel, mu = ee.synthetic.plummer(n)
      for mode, mname in modes:
          single = None
          if rank == 0:
              p = ee.NBodyPropagator.new(ee.Forward(h), 0.0, pos, vel, mu, mode=mode, device=local)
              p.step(steps)
              single = p.state()
              p.close()
          for exch, ename in ((ee.EXCHANGE_ALLGATHER, "allgather"), (ee.EXCHANGE_ALLREDUCE, "allreduce"), (ee.EXCHANGE_ALLREDUCE, "p2p")):
              if mode == ee.MODE_PARITY and exch == ee.EXCHANGE_ALLREDUCE:
                  continue  # unsupported by design (summation order would change): EE_ERR_UNSUPPORTED
              if ename == "p2p" and n < 32768:
                  continue  # the NVLink peer path rides on the pair-symmetric kernel
              uid = eed.broadcast_unique_id(dist, ee.nccl_unique_id() if rank == 0 else None, device="cuda")
              p = ee.NBodyPropagator.new(ee.Forward(h), 0.0, pos, vel, mu, mode=mode, device=local, rank=rank, world=world,
                                         unique_id=uid, exchange=exch)
              if ename == "p2p":
                  eed.connect_peers(dist, p, device="cuda")
              p.step(steps)
              t, y, dy = p.state()
              p.close()
              if rank == 0:
                  st, sy, sdy = single
                  rel = float(np.max(np.linalg.norm(y - sy, axis=1) / np.linalg.norm(sy, axis=1)))
                  relv = float(np.max(np.linalg.norm(dy - sdy, axis=1) / np.linalg.norm(sdy, axis=1)))
                  bit = bool(np.array_equal(y.view(np.uint64), sy.view(np.uint64)) and np.array_equal(dy.view(np.uint64), sdy.view(np.uint64)))
                  # allgather matches bitwise while both runs use the plain kernel; the peer path adds partials in canonical item order,
                  # so it matches the 1-GPU pair-symmetric run bitwise (start-up steps go through NCCL allreduce, so only
                  # positions/velocities produced by the steady steps are expected to agree to the last bit when steps > 12)
                  # (throughput mode on one GPU runs the pair-symmetric kernel from 2 048 bodies up, the target-sharded run the
                  # plain kernel: bitwise agreement is a parity-mode property)
                  expect_bit = ename == "allgather" and mode == ee.MODE_PARITY
                  good = (bit if expect_bit else rel <= 1e-12) and t == st
                  ok = ok and good
                  print(json.dumps({"check": "sharded_vs_single", "world": world, "n": n, "mode": mname, "exchange": ename, "bitwise": bit,
                                    "rel_pos": rel, "rel_vel": relv, "ok": good}), flush=True)
              dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
