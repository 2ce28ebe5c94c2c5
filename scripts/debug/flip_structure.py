"""Debug: structure of the TF32 m=1000 mismatch (is it one flipped select?)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import descent_b200 as d
from helpers import SEED_BASE, init_example_params, synthetic_batch, upload, max_rel_err
from oracle import cpu_ref

m = 1000
env = d.Environment(0)
env.set_tf32(True)
ex = env.example("conv-net", m, optimizer="descent")
rng = np.random.default_rng(SEED_BASE + m)
params = init_example_params(ex, rng)
params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
clusters = ex.train_graph.export_json()["clusters"]
nodes = set()
for t in env.profile(ex.train_graph, 0, 1):
    if t["label"].startswith("TensorCore"):
        nodes.update(clusters[t["cluster"]]["members"])
upload(env, params)
seed = int(rng.integers(0, 2 ** 32))
env.run(ex.train_graph, seed)
for margin in (2e-7, 1e-6, 4e-6):
    want, band, ties = cpu_ref.check_graph_with_tie_band(ex.train_graph_json, params, seed, tf32_nodes=nodes, margin=margin)
    print("margin", margin, "ties", ties)
    for pid, w in want.items():
        got = env.read(env.parameter(pid))
        scale = max(float(np.abs(w).max()), 1e-30)
        print("   %s#%d err %.3g band %.3g" % (env.parameter(pid).name(), pid, np.abs(got - w).max() / scale, band[pid] / scale))
want = cpu_ref.check_graph(ex.train_graph_json, params, seed, tf32_nodes=nodes)
strict = cpu_ref.check_graph(ex.train_graph_json, params, seed)
for p in ex.parameters:
    if p.shape() == [1568, 128]:
        got = env.read(p).astype(np.float64)
        err = np.abs(got - want[p.id])
        scale = np.abs(want[p.id] - params[p.id]).max()
        cols = np.where(err.max(0) > 1e-4 * scale)[0]
        rows = np.where(err.max(1) > 1e-4 * scale)[0]
        print("fc1 w: update scale %.3g, bad cols %s, bad rows %d" % (scale, cols, len(rows)))
        u, s, vt = np.linalg.svd(got - want[p.id])
        print("singular values of the error:", s[:5])
        print("tf32 oracle vs strict oracle:", max_rel_err(want[p.id], strict[p.id]), "gpu vs strict", max_rel_err(got, strict[p.id]))
env.close()
