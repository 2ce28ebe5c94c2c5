"""Parity in the regime the benchmark runs ("what's weak" 1; ADVICE codegen.cpp:930).

bench.py runs conv-net at m = 8192: every persistent kernel (halo conv forward / backward-input / weight gradient,
thin-rows and thin-reduce GEMMs, k-sliced reduces) walks 14-28 tiles per CTA there -- mbarrier phase flips, the
staged-epilogue buffer reuse, the next-halo prefetch, TMEM accumulators carried across tiles, per-CTA partials and
their split sums.  Two ways to reach that code with a checker beside it:

  * `Environment.set_sm_count(2)` plans a graph as if the device had 2 SMs, so the small shapes of the other parity
    tests become many-tiles-per-CTA launches (same kernels, same loops, grid of 2 x resident CTAs);
  * the benchmark's own sizes -- one SGD step of conv-net at m = 1000 (the reference's default, main.rs:106) and
    m = 8192 (bench.py) -- against oracle.cpu_ref in checker mode (float64-accumulated sums like oracle.interp, which
    tests/test_oracle_kat.py::test_cpu_ref_checker_mode_matches_numpy_oracle pins to the numpy interpreter).

Tolerances are BASELINE.json's: 1e-5 of each tensor's maximum for the strict-FP32 path; for TF32 operands 1e-4 against
the oracle that truncates the operands of exactly the MatMuls that ran on tensor cores (as in test_gpu_tf32_and_dp.py)."""
import numpy as np
import pytest

import descent_b200 as d
from helpers import SEED_BASE, init_example_params, max_rel_err, synthetic_batch, upload
from oracle import cpu_ref, run_graph
from test_gpu_conv_kernels import FP32_TOL, SHAPES, TF32_TOL

pytestmark = pytest.mark.gpu


def tensor_core_nodes(env, ex):
    clusters = ex.train_graph.export_json()["clusters"]
    nodes, labels = set(), []
    for t in env.profile(ex.train_graph, 0, 1):
        labels.append(t["label"])
        if t["label"].startswith("TensorCore"):
            nodes.update(clusters[t["cluster"]]["members"])
    return nodes, labels


@pytest.mark.parametrize("tf32", [True, False], ids=["tf32", "strict"])
@pytest.mark.parametrize("shape", SHAPES[:5] + [(64, 14, 14, 16, 32, 2), (40, 28, 28, 1, 16, 1)], ids=lambda s: "x".join(map(str, s)))
def test_conv2d_step_many_tiles_per_cta(env, shape, tf32):
    """test_gpu_conv_kernels.py's step with the planner told the device has 2 SMs: grids shrink to 2 x resident CTAs, so
    each CTA of the halo kernels walks up to dozens of tiles (64 images of 14x14 = 128 tiles over <= 16 CTAs)."""
    m, hh, ww, ic, oc, groups = shape
    env.set_tf32(tf32)
    env.set_sm_count(2)
    x = env.trainable_parameter([m, hh, ww, ic], "x")
    w_out = env.static_parameter([m, hh, ww, oc], "r")
    conv = d.Conv2D(env, ic, oc, 3, 3, pad=1, stride=(1, 1), groups=groups)
    scope = env.scope()
    y = conv.train(scope.parameter(x))
    (y * y * scope.parameter(w_out)).reduce_sum(-1, True).reduce_sum(-2, True).reduce_sum(-3, True).reshape([m, 1]).set_loss()
    ps = scope.trainable_parameters()
    d.StochasticGradientDescent(env, scope, ps, 1.0, 0.0)
    graph_json = scope.export_json()
    graph = scope.build_graph()
    rng = np.random.default_rng(hash(shape) % 1000 + 1)
    vals = {p.id: rng.standard_normal(p.shape()).astype(np.float32) for p in ps}
    vals[w_out.id] = rng.standard_normal(w_out.shape()).astype(np.float32)
    upload(env, vals)
    env.run(graph, 1)
    grids = {t["label"]: t["grid"][0] for t in env.profile(graph, 0, 1) if t["label"].startswith("TensorCore")}
    upload(env, vals)  # the profiling pass ran the step
    env.run(graph, 1)
    want = run_graph(graph_json, vals, 1)
    tol = TF32_TOL if tf32 else FP32_TOL
    errs = {}
    for p in ps:
        grad_got = vals[p.id].astype(np.float64) - env.read(p).astype(np.float64)
        grad_want = vals[p.id].astype(np.float64) - want[p.id].astype(np.float64)
        errs[p.name()] = max_rel_err(grad_got, grad_want)
    print(shape, "tf32" if tf32 else "strict", errs, grids)
    assert max(errs.values()) <= tol, errs
    if tf32 and shape == (64, 14, 14, 16, 32, 2):
        assert grids and max(grids.values()) <= 16, grids  # 128+ tiles over at most 2 x 8 CTAs: >= 8 tiles per CTA


@pytest.mark.parametrize("precision", ["strict", "tf32"])
@pytest.mark.parametrize("network,m", [("conv-net", 64), ("conv-net", 37), ("conv-blur-net", 32), ("single-layer", 512), ("multi-hash", 2048)])
def test_network_step_many_tiles_per_cta(env, network, m, precision):
    """Whole training steps (SGD; Adam for image_fit) planned for 2 SMs, against the oracle with TF32 truncation on the
    MatMuls that ran on tensor cores."""
    optimizer = "adam" if network == "multi-hash" else "descent"
    env.set_tf32(precision == "tf32")
    env.set_sm_count(2)
    ex = env.example(network, m, optimizer=optimizer)
    rng = np.random.default_rng(SEED_BASE + 200 + m)
    params = init_example_params(ex, rng)
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    nodes, _ = tensor_core_nodes(env, ex) if precision == "tf32" else (set(), [])
    upload(env, params)
    from helpers import fill_missing_inputs
    fill_missing_inputs(env, ex.train_graph_json, params)
    seed = int(rng.integers(0, 2 ** 32))
    env.run(ex.train_graph, seed)
    want = run_graph(ex.train_graph_json, params, seed, tf32=("trunc", nodes) if nodes else None)
    tol = 1e-5 if precision == "strict" else (3e-4 if network == "multi-hash" else 1e-4)
    theta = {p.id for p in ex.parameters} if optimizer == "adam" else set()  # Adam's first step: see test_gpu_networks.check_adam_update
    worst = {pid: max_rel_err(env.read(env.parameter(pid)), w) for pid, w in want.items() if pid not in theta}
    print(network, m, precision, worst)
    assert max(worst.values()) <= tol, worst


@pytest.mark.parametrize("precision", ["strict", "tf32"])
@pytest.mark.parametrize("m", [1000, 8192])
def test_conv_net_step_at_benchmark_batch(env, m, precision):
    """bench.py's workload at the reference's default batch and at the benchmark's: one SGD step of conv-net, every output
    (loss / accuracy sums, all 8 parameter tensors = all gradients) against oracle.cpu_ref in checker mode."""
    env.set_tf32(precision == "tf32")
    ex = env.example("conv-net", m, optimizer="descent")
    rng = np.random.default_rng(SEED_BASE + m)
    params = init_example_params(ex, rng)
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    nodes, labels = tensor_core_nodes(env, ex) if precision == "tf32" else (set(), [])
    if precision == "tf32":
        assert sum(l.startswith("TensorCore") for l in labels) >= 6, labels  # conv2 fwd/dF/dX + dense layers
    upload(env, params)
    seed = int(rng.integers(0, 2 ** 32))
    env.run(ex.train_graph, seed)
    want = cpu_ref.check_graph(ex.train_graph_json, params, seed, tf32_nodes=nodes)
    tol = 1e-5 if precision == "strict" else 1e-4
    worst = {env.parameter(pid).name() + "#%d" % pid: max_rel_err(env.read(env.parameter(pid)), w) for pid, w in want.items()}
    print("conv-net m=%d %s:" % (m, precision), worst)
    assert max(worst.values()) <= tol, worst
    # gradients, not just parameters: the update theta' - theta = -lr * g is 2-10 % of max|theta| for the weight tensors
    # (so the 1e-5 above already bounds it at <= 6e-4) and IS the parameter for the zero-initialised biases (1e-5 direct)
    for p in ex.parameters:
        upd_got = env.read(p).astype(np.float64) - params[p.id]
        upd_want = want[p.id].astype(np.float64) - params[p.id]
        assert max_rel_err(upd_got, upd_want) <= (1e-3 if precision == "strict" else 5e-3), (p.name(), p.id, max_rel_err(upd_got, upd_want))
