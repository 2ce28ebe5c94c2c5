"""Calibration run for tests/test_gpu_tf32_and_dp.py::test_tf32_conv_net_100_steps_within_stated_tolerance: N steps of
conv-net on the GPU (strict FP32 and TF32 operands) against the strict CPU oracle, for SGD and Adam; prints the drifts the
test's tolerances are taken from.  python scripts/debug/calibrate_tf32_contract.py [steps] [m]"""
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
    for optimizer in ("descent", "adam"):
        envs = {}
        exs = {}
        for mode in ("strict", "tf32"):
            envs[mode] = d.Environment(0)
            envs[mode].set_tf32(mode == "tf32")
            exs[mode] = envs[mode].example("conv-net", m, optimizer=optimizer)
        ex = exs["strict"]
        rng = np.random.default_rng(77)
        params = init_example_params(ex, rng)
        for mode in envs:
            upload(envs[mode], params)
        program = cpu_ref.Program(ex.train_graph_json)
        state = {pid: np.ascontiguousarray(v, np.float32) for pid, v in params.items()}
        for step in range(steps):
            x, y = synthetic_batch(ex, rng)
            seed = int(rng.integers(0, 2 ** 32))
            for mode in envs:
                envs[mode].write(exs[mode].x, x)
                envs[mode].write(exs[mode].y, y)
                envs[mode].run(exs[mode].train_graph, seed)
            state[ex.x.id], state[ex.y.id] = x, y
            out, _ = program.run(state, seed)
            state.update({pid: v.copy() for pid, v in out.items()})
        program.close()
        want = float(state[ex.loss_sum.id].reshape(-1)[0])
        acc_want = float(state[ex.accuracy_sum.id].reshape(-1)[0])
        for mode in envs:
            env, e = envs[mode], exs[mode]
            got = env.read_parameter_scalar(e.loss_sum)
            acc = env.read_parameter_scalar(e.accuracy_sum)
            drift = {p.name() + "#%d" % p.id: "%.3g" % max_rel_err(env.read(p), state[p.id]) for p in e.parameters}
            rms = {p.name() + "#%d" % p.id: "%.3g" % float(np.linalg.norm(env.read(p).astype(np.float64) - state[p.id]) / np.linalg.norm(state[p.id].astype(np.float64)))
                   for p in e.parameters}
            print("%s %s %d steps m=%d: loss %.6g vs %.6g (rel %.3g) accuracy %g vs %g\n   max/max|theta| %s\n   rms %s" %
                  (optimizer, mode, steps, m, got, want, abs(got - want) / abs(want), acc, acc_want, drift, rms), flush=True)


if __name__ == "__main__":
    main()
