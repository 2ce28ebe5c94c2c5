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


# Near-tie band.  Two correct FP32 evaluations of one step differ in the last bits of every sum (order of additions), so a
# CompareAndSelect whose operands are closer than that noise -- leaky_relu's `x > 0` for a pre-activation within ~1e-7 of
# zero, max_pool2d's `a == max` for a runner-up within ~1e-7 of the winner -- resolves either way, and each such flip moves a
# mean-over-the-batch gradient by one sample's full contribution (measured: up to 3e-2 of a bias gradient at m = 1000).  At
# mini-batches of thousands of samples (2e7 ... 1.6e8 select elements per step) a few flips always happen; small batches
# (every other parity test) rarely see one.  oracle.cpu_ref.check_graph_with_tie_band evaluates the step three times --
# as computed, and with every order-sensitive select that is within `margin` of a tie forced true resp. false -- and
# returns the band those resolutions span per output.  The assertion is  |gpu - oracle| <= tol * max|oracle| + 3 * band.
# margin = the forward noise between the two evaluations: 2e-7 of the tensor's maximum for strict FP32; 1e-5 for TF32
# operands, where a last-bit difference in one layer's output can cross a truncation boundary of the next layer's
# operand (2^-11 of that element; measured 2.4e-6 of max|z| at the first dense layer).
TIE_MARGIN = {"strict": 2e-7, "tf32": 1e-5}


def tensor_core_nodes(env, ex):
    clusters = ex.train_graph.export_json()["clusters"]
    nodes, labels = set(), []
    for t in env.profile(ex.train_graph, 0, 1):
        labels.append(t["label"])
        if t["label"].startswith("TensorCore"):
            for ci in t["clusters"]:
                nodes.update(clusters[ci]["members"])
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
@pytest.mark.parametrize("network,m", [("conv-net", 64), ("conv-net", 37), ("conv-blur-net", 32), ("single-layer", 512), ("multi-hash", 2048), ("multi-hash", 2085), ("multi-hash", 4096), ("multi-hash", 4133)])
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
    nodes, labels = tensor_core_nodes(env, ex) if precision == "tf32" else (set(), [])
    if network == "multi-hash" and m >= 4096 and precision == "tf32":  # the fused MLP kernel (dense chain) and the grouped scatter are what runs
        assert any("DenseChain" in l for l in labels) and any("ScatterAdd group" in l for l in labels), labels
    upload(env, params)
    from helpers import fill_missing_inputs
    fill_missing_inputs(env, ex.train_graph_json, params)
    seed = int(rng.integers(0, 2 ** 32))
    env.run(ex.train_graph, seed)
    want, band, ties = cpu_ref.check_graph_with_tie_band(ex.train_graph_json, params, seed, tf32_nodes=nodes, margin=TIE_MARGIN[precision])
    tol = 1e-5 if precision == "strict" else (3e-4 if network == "multi-hash" else 1e-4)
    theta = {p.id for p in ex.parameters} if optimizer == "adam" else set()  # Adam's first step: see test_gpu_networks.check_adam_update
    worst, failed = {}, {}
    for pid, w in want.items():
        if pid in theta:
            continue
        err = max_rel_err(env.read(env.parameter(pid)), w)
        worst[pid] = err
        if not err <= tol + 3.0 * band[pid] / max(float(np.abs(w).max()), 1e-30):  # near-tie band: see TIE_MARGIN below
            failed[pid] = err
    print(network, m, precision, "ties", ties, worst)
    assert not failed, (failed, worst)




@pytest.mark.parametrize("precision", ["strict", "tf32"])
@pytest.mark.parametrize("m", [1000, 8192])
def test_conv_net_step_at_benchmark_batch(env, m, precision):
    """bench.py's workload at the reference's default batch and at the benchmark's: one SGD step of conv-net, every output
    (loss / accuracy sums, momentum state = gradients, all 8 parameter tensors) against oracle.cpu_ref in checker mode.
    Loss and accuracy sums do not depend on any select resolution: 1e-5 / exact in both precisions."""
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
    want, band, ties = cpu_ref.check_graph_with_tie_band(ex.train_graph_json, params, seed, tf32_nodes=nodes, margin=TIE_MARGIN[precision])
    tol = 1e-5 if precision == "strict" else 1e-4
    report, failed = {}, {}
    for pid, w in want.items():
        name = env.parameter(pid).name() + "#%d" % pid
        scale = max(float(np.abs(w).max()), 1e-30)
        err = max_rel_err(env.read(env.parameter(pid)), w)
        allowed = tol + 3.0 * band[pid] / scale
        report[name] = "%.2g (band %.2g)" % (err, band[pid] / scale)
        if not err <= allowed:
            failed[name] = (err, allowed)
    print("conv-net m=%d %s: %d near-tie select elements at margin %g; error vs oracle (tie band):" % (m, precision, ties, TIE_MARGIN[precision]), report)
    assert not failed, failed
    assert max_rel_err(env.read(ex.loss_sum), want[ex.loss_sum.id]) <= 1e-5
    np.testing.assert_array_equal(env.read(ex.accuracy_sum), want[ex.accuracy_sum.id])


def test_options_changed_after_a_run_replan_the_graph(env):
    """set_tf32 / set_sm_count after a graph has run used to be ignored silently (the plan was cached).  The
    executor now plans the graph again under the new options: the kernel set changes, and both results match their oracle."""
    ex = env.example("conv-net", 32, optimizer="descent")
    rng = np.random.default_rng(3)
    params = init_example_params(ex, rng)
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    upload(env, params)
    env.run(ex.train_graph, 9)
    strict_labels = [t["label"] for t in env.profile(ex.train_graph, 0, 1)]
    assert not any(l.startswith("TensorCore") for l in strict_labels)
    env.set_tf32(True)
    nodes, tf32_labels = tensor_core_nodes(env, ex)
    assert any(l.startswith("TensorCore") for l in tf32_labels), tf32_labels
    env.set_sm_count(3)
    grids = {t["label"]: t["grid"][0] for t in env.profile(ex.train_graph, 0, 1) if t["label"].startswith("TensorCore") and t["grid"][0] > 1}
    assert grids and max(grids.values()) <= 3 * 8, grids  # persistent grids follow the overridden SM count
    upload(env, params)
    env.run(ex.train_graph, 9)
    want = run_graph(ex.train_graph_json, params, 9, tf32=("trunc", nodes))
    worst = {pid: max_rel_err(env.read(env.parameter(pid)), w) for pid, w in want.items()}
    assert max(worst.values()) <= 1e-4, worst
