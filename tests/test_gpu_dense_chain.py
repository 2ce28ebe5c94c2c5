"""The fused MLP training-step kernel (graph.hpp DenseChain, codegen generate_dense_chain_code) on networks other than the
image_fit head it was built for: input widths that are not multiples of four (scalar x_0 staging, zero-padded k), no
gradient with respect to the input, hidden widths 16 / 32 / 128, a five-wide output, ragged batches -- each one Adam step
against the oracle with TF32 truncation on the MatMuls the kernel covers (tolerance as in test_gpu_tf32_and_dp.py)."""
import numpy as np
import pytest

import descent_b200 as d
from helpers import max_rel_err
from oracle import run_graph

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("widths,m", [((6, 32, 16, 5), 4115), ((20, 64, 64, 3), 4096), ((12, 64, 32, 8), 8192 + 37), ((8, 16, 16, 16, 4), 4100)], ids=str)
def test_custom_mlp_step_runs_as_one_kernel_and_matches_the_oracle(env, widths, m):
    env.set_tf32(True)
    env.set_sm_count(3)  # several tiles per CTA: the weight gradients accumulate in TMEM across tiles
    x = env.static_parameter([m, widths[0]], "x")
    y = env.static_parameter([m, widths[-1]], "y")
    loss_sum = env.static_parameter([1], "loss")
    layers = [d.Dense(env, a, b) for a, b in zip(widths[:-1], widths[1:])]
    scope = env.scope()
    h = scope.parameter(x)
    for i, layer in enumerate(layers):
        h = layer.train(h)
        if i + 1 < len(layers):
            h = h.leaky_relu(0.01)
    loss = (h - y).square().reduce_sum(-1, True).set_loss()
    scope.update_parameter_value(loss_sum, lambda s: s + loss.reduce_sum(0, False))
    params = scope.trainable_parameters()
    opt = d.Adam(env, scope, params, 0.01, 0.9, 0.99, 1.0e-8)
    graph_json = scope.export_json()
    graph = scope.build_graph()
    exported = graph.export_json()
    assert len(exported["dense_chains"]) == 1 and exported["dense_chains"][0]["widths"] == list(widths), exported["dense_chains"]
    rng = np.random.default_rng(sum(widths) + m)
    values = {p.id: (rng.standard_normal(p.shape()) * (0.5 if len(p.shape()) == 2 else 0.1)).astype(np.float32) for p in params}
    for p in opt.state():
        values[p.id] = np.zeros(p.shape(), np.float32)
    values[x.id] = rng.standard_normal((m, widths[0])).astype(np.float32)
    values[y.id] = rng.standard_normal((m, widths[-1])).astype(np.float32)
    values[loss_sum.id] = np.zeros(1, np.float32)

    def upload():
        for pid, v in values.items():
            env.write(env.parameter(pid), v)
    upload()
    launches = env.profile(graph, 0, 1)
    labels = [t["label"] for t in launches]
    assert sum("DenseChain" in l for l in labels) == 1 and not any(l.startswith("TensorCoreMatMul") or l.startswith("MatMul") for l in labels), labels
    nodes = set()
    for t in launches:
        if t["label"].startswith("TensorCore"):
            for ci in t["clusters"]:
                nodes.update(exported["clusters"][ci]["members"])
    upload()
    env.run(graph, 3)
    want = run_graph(graph_json, values, 3, tf32=("trunc", nodes))
    theta = {p.id for p in params}  # Adam's first step is sign-like in near-zero gradients: compared through m and v (test_gpu_networks.py)
    worst = {}
    for pid, w in want.items():
        if pid not in theta:
            worst[env.parameter(pid).name() + "#%d" % pid] = max_rel_err(env.read(env.parameter(pid)), w)
    print(widths, m, worst)
    assert max(worst.values()) <= 1e-4, worst
