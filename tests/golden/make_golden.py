"""Generates tests/golden/*.npz: seeded inputs and the oracle's outputs for one SGD training step of each example
network at a tiny batch, plus the integer vectors of SURVEY.md Appendix D (pcg hash, Rand, hash-grid indices).

    python tests/golden/make_golden.py          (CPU only; needs the built library for the host-only graph builder)

The reference itself cannot run in this image (Rust + Vulkan), so these fixtures pin the
ORACLE: `tests/test_oracle_kat.py::test_oracle_reproduces_golden_steps` fails if its arithmetic ever drifts, and
`tests/test_gpu_networks.py::test_cuda_matches_golden_steps` compares the CUDA backend with the committed numbers
without executing the oracle.  The oracle in turn is pinned by the reference's own tests (tests/reference_kats.py).
SGD is used because its outputs (theta - lr * gradient) are well conditioned in every entry, unlike Adam's first step.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import descent_b200 as d  # noqa: E402
from helpers import init_example_params, synthetic_batch  # noqa: E402
from oracle import interp, run_graph  # noqa: E402

STEPS = [("linear", 8), ("single-layer-dropout", 8), ("conv-net", 4), ("conv-blur-net", 2)]
SEED = 20260117


def example_step(env, network, m):
    ex = env.example(network, m, optimizer="descent")
    rng = np.random.default_rng(SEED + len(network))
    params = init_example_params(ex, rng)
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    seed = int(rng.integers(0, 2 ** 32))
    return ex, params, seed


def main():
    for network, m in STEPS:
        env = d.Environment(-1)  # fresh parameter ids per network, as in the tests
        ex, params, seed = example_step(env, network, m)
        missing = [n["parameter"] for n in ex.train_graph_json["nodes"] if n["op"] == "Input" and n["parameter"] not in params]
        assert not missing or network == "conv-blur-net", missing
        for pid in missing:  # MaxBlurPool2D's fixed depthwise filter [1, 2, 1] x [1, 2, 1] / 16 and its zero bias (module.rs:139-163)
            shape = env.parameter(pid).shape()
            blur = np.outer([1, 2, 1], [1, 2, 1]).astype(np.float32) / 16.0
            params[pid] = np.broadcast_to(blur.reshape(1, 1, 3, 3, 1), shape).astype(np.float32).copy() if len(shape) == 5 else np.zeros(shape, np.float32)
        want = run_graph(ex.train_graph_json, params, seed)
        arrays = {"seed": np.array([seed], np.uint64)}
        for pid, v in params.items():
            arrays["in_%d" % pid] = v
        for pid, v in want.items():
            arrays["out_%d" % pid] = v
        np.savez_compressed(os.path.join(HERE, "step_%s_m%d.npz" % (network, m)), **arrays)
        print(network, m, "inputs", len(params), "outputs", len(want))
    idx = np.arange(4096, dtype=np.uint32)
    np.savez_compressed(os.path.join(HERE, "integer_vectors.npz"), index=idx, pcg=interp.pcg(idx),
                        rand_uid3_seed7=interp.rand_from_index(3, idx, 7).view(np.uint32))


if __name__ == "__main__":
    main()
