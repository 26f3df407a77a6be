"""BASELINE config c3 (or c5) at full size: `--utterances` synthetic 3-second utterances, utterance-sharded over the ranks of
a torchrun launch (round-robin, no data-path collective), each rank through corpus.analyze_corpus (host synthesis of the
inputs in threads ahead of the GPU, memory-sized batches, results copied to the host).  Wall time from the first batch to
the last result, max over ranks; one JSON line from rank 0.

    torchrun --nproc-per-node 8 tools/run_corpus.py --utterances 10000 [--config c3|c5] [--workers 8]
"""
import argparse, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from eaqhm_b200 import corpus, engine, host, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--utterances", type=int, default=10000)
ap.add_argument("--config", default="c3", choices=["c3", "c5"])
ap.add_argument("--workers", type=int, default=8)
ap.add_argument("--batch-gb", type=float, default=24.0)
ap.add_argument("--preload", action="store_true",
                help="synthesise the rank's utterances before the clock starts (a corpus on disk: inputs are host arrays already)")
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))


def load(i):
    if args.config == "c3":
        return engine.Utterance(**bench.corpus_inputs(i, 1)[0])
    cfg = synthetic.config_c5(i, duration=3.0)
    s, f0s = synthetic.make(cfg)
    t5 = (np.arange(1, len(s), 80) - 1) / 16000
    v5 = np.asarray(cfg["voiced_mask_of_t"](t5), bool)
    return engine.Utterance(s=s, fs=16000, f0s=f0s, voiced5=v5, frame_step=80, f0min=250.0,
                            Kmax=host.harmonic_capacity(16000, f0s)[1], det_part=host.voiced_only_target(s, v5, v5, 80))


synth = load
if args.preload:
    from concurrent.futures import ThreadPoolExecutor
    from eaqhm_b200.sharding import partition_round_robin
    mine = partition_round_robin(args.utterances, rank, world)
    with ThreadPoolExecutor(max_workers=args.workers) as pool:
        cache = dict(zip(mine, pool.map(synth, mine)))
    load = lambda i: cache[i] if i in cache else synth(i)

# warm-up: library load, workspace, pinned staging
for _ in corpus.analyze_corpus(min(8, args.utterances), load, rank=0, world=1, workers=args.workers):
    pass
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
n, audio, solves, adapt, srer, not_pd = 0, 0.0, 0, [], [], 0
for i, r in corpus.analyze_corpus(args.utterances, load, rank=rank, world=world, batch_bytes=int(args.batch_gb * (1 << 30)),
                                  workers=args.workers, lookahead=256):
    n += 1
    audio += len(r.s_recon) / r.fs
    solves += int(r.isVoiced.sum()) * len(r.SRER)
    adapt.append(len(r.SRER))
    srer.append(max(r.SRER))
    not_pd += r.stats["not_pd_frames"]
torch.cuda.synchronize()
dt = time.perf_counter() - t0
row = torch.tensor([dt, audio, solves, n, float(np.sum(adapt)), float(np.sum(srer)), not_pd], dtype=torch.float64, device="cuda")
if dist is not None:
    rows = [torch.empty_like(row) for _ in range(world)]
    dist.all_gather(rows, row)
    rows = torch.stack(rows).cpu().numpy()
else:
    rows = row.cpu().numpy()[None]
if rank == 0:
    wall = float(rows[:, 0].max())
    tot_audio, tot_solves, tot_n = float(rows[:, 1].sum()), float(rows[:, 2].sum()), int(rows[:, 3].sum())
    print(json.dumps({"workload": "%s: %d synthetic 3-second utterances at 16 kHz, utterance-sharded over %d GPU(s), corpus runner "
                                  "(%s, results on the host)"
                                  % (args.config, tot_n, world, "inputs as host arrays before the clock starts" if args.preload else
                                     "host synthesis of the inputs in %d threads per rank ahead of the GPU" % args.workers),
                      "n_gpus": world, "wall_seconds_max_over_ranks": wall, "wall_seconds_per_rank": [float(x) for x in rows[:, 0]],
                      "audio_seconds": tot_audio, "audio_s_per_s": tot_audio / wall, "frame_solves": int(tot_solves),
                      "frame_solves_per_s": tot_solves / wall, "mean_adaptations": float(rows[:, 4].sum() / tot_n),
                      "mean_best_SRER_dB": float(rows[:, 5].sum() / tot_n), "not_pd_frames": int(rows[:, 6].sum())}))
if dist is not None:
    dist.destroy_process_group()
