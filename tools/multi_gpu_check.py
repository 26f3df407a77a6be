"""torchrun check of the frame-range sharded mode over NCCL: every rank streams its contiguous block of shards of one
signal, the per-frame parameters of the boundary frames are exchanged with the neighbours after every analysis pass
(NCCL send/recv), the SRER statistics are all-gathered; rank 0 compares the assembled result with the monolithic
single-GPU run.  Usage: torchrun --nproc-per-node N tools/multi_gpu_check.py [seconds] [shard_seconds] [maxAdpt]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
from eaqhm_b200 import engine, longsignal, synthetic, host

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
shard_s = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
maxAdpt = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = synthetic.config_c2(duration=seconds)
s, f0s = synthetic.make(cfg)
n5 = len(np.arange(1, len(s), 80))
Kmax = host.harmonic_capacity(16000, f0s)[1]
res, plan, own = longsignal.analyze_long_signal(s, 16000, f0s, np.ones(n5, bool), 80, 90.0, Kmax, shard_seconds=shard_s,
                                                maxAdpt=maxAdpt, rank=rank, world=world, dist=dist)
parts = [None] * world
dist.all_gather_object(parts, (own, res.s_recon, res.am, res.fm, res.SRER, res.stats))
if rank == 0:
    u = engine.Utterance(s=s, fs=16000, f0s=f0s, voiced5=np.ones(n5, bool), frame_step=80, f0min=90.0, Kmax=Kmax)
    mono = engine.Engine([u], maxAdpt=maxAdpt).run()[0]
    s_all = sum(p[1] for p in parts); am_all = sum(p[2] for p in parts); fm_all = sum(p[3] for p in parts)
    same = all(p[4] == parts[0][4] for p in parts)
    print("owned ranges", [p[0] for p in parts], "shards per rank", [p[5]["shards"] for p in parts])
    print("SRER sharded", parts[0][4], "\nSRER mono   ", mono.SRER, "\nsame SRER list on all ranks:", same)
    print("halo bytes sent per rank", [p[5]["exchanged_bytes"] for p in parts],
          "frame solves", sum(p[5]["frames_solved"] for p in parts), "mono", mono.stats["frames_solved"])
    ds, dam = np.abs(s_all - mono.s_recon).max(), np.abs(am_all - mono.am).max()
    dfm = np.abs(fm_all - mono.fm).max()
    print("max|ds|", ds, "max|dam|", dam, "max|dfm|", dfm)
    ok = (ds < 1e-9 and same and len(parts[0][4]) == len(mono.SRER)
          and np.abs(np.array(parts[0][4]) - np.array(mono.SRER)).max() < 1e-9
          and sum(p[5]["frames_solved"] for p in parts) == mono.stats["frames_solved"])
    print("MULTI_GPU_CHECK", "OK" if ok else "FAIL")
dist.destroy_process_group()
