"""torchrun --nproc-per-node N tools/e2e_breakdown_multi.py : per-rank wall time of the three parts of the end-to-end call
(engine construction from host arrays, device loop with prefetched read-back, results) and a cProfile of construction and
results on rank 0."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
import bench
from eaqhm_b200 import api

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
inp = bench.long_signal_inputs(20.0 * world)
ones = np.ones(len(inp["voiced5"]), bool)
make = lambda: api.make_engine(inp["s"], inp["fs"], inp["f0s"], ones, ones, inp["frame_step"], gender=(90, 250),
                               shard_seconds=20.0, rank=rank, world=world, dist=dist)
for rep in range(5):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    h = make(); h.eng.prepare_readback()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = h.run_device(fixed_adaptations=6, prefetch_results=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    r = h.results(out)
    t3 = time.perf_counter()
    del h
    print("rank %d rep %d: construct %.1f ms, loop %.1f ms, results %.1f ms" % (rank, rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
if rank == 0:
    pr = cProfile.Profile()
dist.barrier()
if rank == 0:
    pr.enable()
h = make(); torch.cuda.synchronize()
if rank == 0:
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(16)
out = h.run_device(fixed_adaptations=6, prefetch_results=True); torch.cuda.synchronize()
if rank == 0:
    pr = cProfile.Profile(); pr.enable()
r = h.results(out)
if rank == 0:
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
dist.destroy_process_group()
