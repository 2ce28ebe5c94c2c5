"""Debug (torchrun, N ranks): per-launch event timings of the data-parallel conv-net step on rank 0."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import descent_b200 as d
from helpers import init_example_params, synthetic_batch, upload

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
env = d.Environment(local)
uid = [d.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
env.init_data_parallel(world, rank, uid[0])
ready = ctypes.c_int(-1)
d.lib.dsc_dp_peer_memory_ready(env.ctx(), ctypes.byref(ready))
env.set_tf32(True)
ex = env.example("conv-net", 8192)
rng = np.random.default_rng(1)
params = init_example_params(ex, rng)
params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, np.random.default_rng(rank))
upload(env, params)
for s in range(3):
    env.run(ex.train_graph, s)
env.sync(); dist.barrier()
prof = env.profile(ex.train_graph, 1, 10)
env.sync(); dist.barrier()
if rank == 0:
    print("peer memory ready:", ready.value, "launches", len(prof), "event sum %.1f us" % (sum(t["ms"] for t in prof) * 1e3))
    for t in prof:
        if t["ms"] * 1e3 > 7 or "AllReduce" in t["label"]:
            print("  %-12s %7.1f us  %s" % (t["entry"][:12], t["ms"] * 1e3, t["label"][:100]))
env.close()
dist.destroy_process_group()
