"""Developer aid: where the end-to-end time goes with several ranks and pre-sharded input (torchrun).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/dev/e2e_timing_multi.py cfg5"""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ssvio_b200 import ba, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
idt = torch.zeros(ba.SSBA_NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
if rank == 0: idt = torch.tensor(list(ba.nccl_unique_id()), dtype=torch.uint8, device=dev)
dist.broadcast(idt, 0)
g = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg3").shard(rank, world)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rows = []
opt = ba.BundleAdjuster(device_id=local, rank=rank, world_size=world, nccl_id=bytes(idt.cpu().tolist()), presharded=True)
for rep in range(reps):
    opt.drop_structure()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); opt.set_graph(g)
    t1 = time.perf_counter(); opt.initialize_optimization()
    t2 = time.perf_counter(); r = opt.optimize(g.iters)
    t3 = time.perf_counter(); p = opt.poses(); q = opt.points()
    t4 = time.perf_counter()
    rows.append([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0])
m = 1e3 * np.median(np.array(rows[3:]), axis=0)
print(f"rank {rank}/{world}: edges {g.n_edges}  set_graph {m[0]:.3f}  initialize {m[1]:.3f}  optimize {m[2]:.3f}  read-back {m[3]:.3f}  total {m[4]:.3f} ms", flush=True)
opt.close(); dist.barrier(); dist.destroy_process_group()
