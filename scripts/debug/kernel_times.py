"""Debug: event timings of the large kernels of one workload (run under gpurun).  usage: kernel_times.py [workload] [m] [min_us]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import descent_b200 as d
from helpers import init_example_params, synthetic_batch, upload

net = sys.argv[1] if len(sys.argv) > 1 else "conv-net"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
min_us = float(sys.argv[3]) if len(sys.argv) > 3 else 15.0
env = d.Environment(0)
env.set_tf32(True)
ex = env.example(net, m)
rng = np.random.default_rng(1)
params = init_example_params(ex, rng, siren=(net == "siren"))
params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
upload(env, params)
for s in range(3):
    env.run(ex.train_graph, s)
prof = env.profile(ex.train_graph, 1, 10)
total = sum(t["ms"] for t in prof)
print("%s m=%d env %s: %d launches, event sum %.1f us" % (net, m, {k: v for k, v in os.environ.items() if k.startswith("DSC_")}, len(prof), total * 1e3))
for t in prof:
    if t["ms"] * 1e3 >= min_us:
        bw = t["bytes"] / (t["ms"] * 1e-3) / 1e9
        print("  %-10s %7.1f us %6.0f MB %5.0f GB/s (%.2f) grid %s smem %d  %s" % (t["entry"][:10], t["ms"] * 1e3, t["bytes"] / 1e6, bw, bw / 6527.8, t["grid"], t["smem"], t["label"][:70]))
env.close()
