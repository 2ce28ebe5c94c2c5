"""Steps conv-net with SGD on the GPU (strict FP32) and in the CPU oracle side by side and reports, per step, the loss of
the step, the largest parameter magnitude and the first step at which the two disagree.
python scripts/debug/sgd_divergence.py [steps] [m] [optimizer]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import descent_b200 as d  # noqa: E402
from helpers import init_example_params, max_rel_err, synthetic_batch, upload  # noqa: E402
from oracle import cpu_ref  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    optimizer = sys.argv[3] if len(sys.argv) > 3 else "descent"
    env = d.Environment(0)
    env.set_tf32(False)
    ex = env.example("conv-net", m, optimizer=optimizer)
    rng = np.random.default_rng(77)
    params = init_example_params(ex, rng)
    upload(env, params)
    program = cpu_ref.Program(ex.train_graph_json)
    state = {pid: np.ascontiguousarray(v, np.float32) for pid, v in params.items()}
    prev_g = prev_o = 0.0
    for step in range(steps):
        x, y = synthetic_batch(ex, rng)
        seed = int(rng.integers(0, 2 ** 32))
        env.write(ex.x, x)
        env.write(ex.y, y)
        env.run(ex.train_graph, seed)
        state[ex.x.id], state[ex.y.id] = x, y
        out, _ = program.run(state, seed)
        state.update({pid: v.copy() for pid, v in out.items()})
        lg = env.read_parameter_scalar(ex.loss_sum)
        lo = float(state[ex.loss_sum.id].reshape(-1)[0])
        worst = max((max_rel_err(env.read(p), state[p.id]), p.name() + "#%d" % p.id) for p in ex.parameters + ex.optimizer_state)
        mag = max(float(np.abs(state[p.id]).max()) for p in ex.parameters)
        print("step %3d: step loss gpu %.6g oracle %.6g | worst tensor drift %.3g (%s) | max|theta| oracle %.4g" % (step, lg - prev_g, lo - prev_o, worst[0], worst[1], mag), flush=True)
        prev_g, prev_o = lg, lo
        if not np.isfinite(lg) or not np.isfinite(lo):
            for p in ex.parameters + ex.optimizer_state:
                g = env.read(p)
                print("   %s#%d gpu finite %s oracle finite %s max|gpu| %s" % (p.name(), p.id, bool(np.isfinite(g).all()), bool(np.isfinite(state[p.id]).all()), float(np.nanmax(np.abs(g)))))
            break
    program.close()


if __name__ == "__main__":
    main()
