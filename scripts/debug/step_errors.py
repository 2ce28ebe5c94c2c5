"""Debug: per-output error of one conv-net SGD step vs cpu_ref checker, several batch sizes (run under gpurun)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import descent_b200 as d
from helpers import init_example_params, synthetic_batch, upload, max_rel_err, fill_missing_inputs
from oracle import cpu_ref

net = sys.argv[1]
for m in [int(x) for x in sys.argv[2].split(",")]:
    for tf32 in (True, False):
        for sm in ([0, 2] if len(sys.argv) > 3 else [0]):
            env = d.Environment(0)
            env.set_tf32(tf32)
            env.set_sm_count(sm)
            ex = env.example(net, m, optimizer="descent")
            rng = np.random.default_rng(m)
            params = init_example_params(ex, rng)
            params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
            clusters = ex.train_graph.export_json()["clusters"]
            nodes = set()
            prof = env.profile(ex.train_graph, 0, 1)
            for t in prof:
                if t["label"].startswith("TensorCore"):
                    nodes.update(clusters[t["cluster"]]["members"])
            upload(env, params)
            fill_missing_inputs(env, ex.train_graph_json, params)
            env.run(ex.train_graph, 7)
            want = cpu_ref.check_graph(ex.train_graph_json, params, 7, tf32_nodes=nodes if tf32 else ())
            worst = {env.parameter(pid).name() + "#%d" % pid: "%.2g" % max_rel_err(env.read(env.parameter(pid)), w) for pid, w in want.items()}
            print(net, "m=%d tf32=%s sm=%d" % (m, tf32, sm), worst)
            if tf32 and m == 1000:
                for t in prof:
                    print("    ", t["label"][:90], t["entry"], t["grid"])
            env.close()
