"""torchrun check of the frame-range sharded mode over NCCL: every rank analyses its block of shards of one signal,
the SRER statistics are all-gathered, rank 0 compares the stitched result with the monolithic single-GPU run."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
from eaqhm_b200 import engine, sharding, synthetic, host

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = synthetic.config_c2(duration=6.0)
s, f0s = synthetic.make(cfg)
n5 = len(np.arange(1, len(s), 80))
Kmax = host.harmonic_capacity(16000, f0s)[1]
res, plan, blk = sharding.analyze_long_signal(s, 16000, f0s, np.ones(n5, bool), 80, 90.0, Kmax, shard_seconds=1.0,
                                              maxAdpt=2, rank=rank, world=world, dist=dist)
parts = [None] * world
dist.all_gather_object(parts, (blk, res.s_recon, res.am, res.SRER))
if rank == 0:
    u = engine.Utterance(s=s, fs=16000, f0s=f0s, voiced5=np.ones(n5, bool), frame_step=80, f0min=90.0, Kmax=Kmax)
    mono = engine.Engine([u], maxAdpt=2).run()[0]
    s_all = sum(p[1] for p in parts); am_all = sum(p[2] for p in parts)
    print("blocks", [p[0] for p in parts], "SRER shards", parts[0][3], "mono", mono.SRER)
    print("same SRER on all ranks:", all(p[3] == parts[0][3] for p in parts))
    print("max|ds|", np.abs(s_all - mono.s_recon).max(), "max|dam|", np.abs(am_all - mono.am).max())
    ok = np.abs(s_all - mono.s_recon).max() < 1e-9 and all(p[3] == parts[0][3] for p in parts)
    print("MULTI_GPU_CHECK", "OK" if ok else "FAIL")
dist.destroy_process_group()
