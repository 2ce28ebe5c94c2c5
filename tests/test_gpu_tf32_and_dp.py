"""Tensor-core (TF32) training path and the NCCL data-parallel path on real GPUs.

TF32 tolerance (BASELINE.json: "within a stated TF32 tolerance on loss and parameters after N steps").  Two statements:
(1) parity: against the oracle run with TF32-emulated MatMul operands (oracle.interp.tf32_operand, truncation as the
    hardware does -- pinned by test_gpu_gemm_tf32.py::test_tf32_operand_rounding_mode) the tensor-core path must agree
    like a strict path does: gradients within 5e-5 of each tensor's max after one step;
(2) drift: against the strict-FP32 oracle, after N = 20 Adam steps of single-layer at m = 1024 the loss stays within
    TF32_LOSS_TOL and every parameter within TF32_PARAM_TOL * max|theta| (the measured drift is printed)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import init_example_params, max_rel_err, synthetic_batch, upload
from oracle import run_graph

pytestmark = pytest.mark.gpu
TF32_LOSS_TOL = 2e-3  # measured 3.2e-4
TF32_PARAM_TOL = 1.5e-1  # measured 1.4e-2 .. 5.3e-2 (Adam steps of near-zero gradients flip sign)
# conv-net, 100 Adam steps at m = 256 (calibration run on a B200, scripts/debug/calibrate_tf32_contract.py; the test's docstring
# explains the yardstick): strict path rms drift 0.001 .. 0.19 per tensor, TF32 path 0.0027 .. 0.32, ratios 1.05 .. 5.3
CONV_NET_TF32_DRIFT_FACTOR = 4.0
CONV_NET_TF32_DRIFT_FLOOR = 1e-2
CONV_NET_TF32_PARAM_RMS_CAP = 6e-1
CONV_NET_TF32_ACCURACY_TOL = 5e-2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tensor_core_nodes(env, ex):
    """Op-node ids of the MatMuls the executor sent to the tcgen05 kernel (launch label 'TensorCore...')."""
    clusters = ex.train_graph.export_json()["clusters"]
    nodes = set()
    for t in env.profile(ex.train_graph, 0, 1):
        if t["label"].startswith("TensorCore"):
            for ci in t["clusters"]:
                nodes.update(clusters[ci]["members"])
    return nodes


def test_tf32_training_drift_within_stated_tolerance(env):
    env.set_tf32(True)
    ex = env.example("single-layer", 1024)
    labels = [t["label"] for t in env.profile(ex.train_graph, 0, 1)]
    assert sum(l.startswith("TensorCore") for l in labels) >= 2, labels  # fc1 forward and dW1 (N = 10 operands are not TMA-aligned)
    rng = np.random.default_rng(21)
    params = init_example_params(ex, rng)
    upload(env, params)
    state = dict(params)
    for step in range(20):
        x, y = synthetic_batch(ex, rng)
        seed = int(rng.integers(0, 2 ** 32))
        env.write(ex.x, x)
        env.write(ex.y, y)
        env.run(ex.train_graph, seed)
        state[ex.x.id], state[ex.y.id] = x, y
        state.update(run_graph(ex.train_graph_json, state, seed))
    got, want = env.read_parameter_scalar(ex.loss_sum), float(state[ex.loss_sum.id][0])
    drift = {p.name() + "#%d" % p.id: max_rel_err(env.read(p), state[p.id]) for p in ex.parameters}
    print("tf32 drift after 20 steps: loss rel %.3g, parameters %s" % (abs(got - want) / abs(want), drift))
    assert abs(got - want) <= TF32_LOSS_TOL * abs(want), (got, want, drift)
    assert max(drift.values()) <= TF32_PARAM_TOL, drift


def test_tf32_single_step_gradients(env):
    """One step: Adam's m state (= 0.1 * gradient) within 5e-5 of the TF32-emulating oracle, relative to each tensor's max."""
    env.set_tf32(True)
    ex = env.example("single-layer", 512)
    rng = np.random.default_rng(4)
    params = init_example_params(ex, rng)
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    nodes = tensor_core_nodes(env, ex)
    assert len(nodes) >= 2
    upload(env, params)
    env.run(ex.train_graph, 1)
    want = run_graph(ex.train_graph_json, params, 1, tf32=("trunc", nodes))
    strict = run_graph(ex.train_graph_json, params, 1)
    for i, p in enumerate(ex.parameters):
        m_state = ex.optimizer_state[1 + 2 * i]
        got = env.read(m_state)
        print("%s: vs tf32-emulating oracle %.3g, vs strict oracle %.3g" % (p.name(), max_rel_err(got, want[m_state.id]), max_rel_err(got, strict[m_state.id])))
        assert max_rel_err(got, want[m_state.id]) <= 5e-5, p.name()


@pytest.mark.parametrize("workload,m,optimizer,precision,steps", [("conv-net", 16, "adam", "strict", 3), ("multi-hash", 16384, "adam", "tf32", 1)], ids=["conv-net", "multi-hash-fused"])
def test_two_gpu_data_parallel_matches_single_gpu(workload, m, optimizer, precision, steps):
    """torchrun, 2 ranks, NCCL bucket all-reduce inside the captured step: parameters after 3 steps equal the
    1-GPU run on the concatenated batch (same tolerance as the CPU data-parallel test).  multi-hash (fused MLP kernel,
    grouped scatter, MLP gradients in the early bucket): ONE Adam step, compared on the optimiser state -- m and v are the
    all-reduced gradients and their squares, whereas Adam's first parameter step is sign-like in near-zero gradients."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(ROOT, "tests", "dp_worker.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", script, workload, str(m), optimizer, precision, str(steps)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DP_OK" in out.stdout, out.stdout[-3000:]
    print(out.stdout[-600:])
    if workload == "multi-hash":  # the fused MLP kernel and the grouped scatter run on every rank, the MLP gradients in the early bucket
        assert "DenseChain" in out.stdout and "ScatterAdd group" in out.stdout and "early bucket" in out.stdout, out.stdout[-3000:]


# SIREN's first layer (x30 init) feeds sin() arguments of magnitude ~50: FP32 accumulation-order noise of the GEMM is
# amplified by the oscillation, hence the wider bound there.
@pytest.mark.parametrize("network,m,tol", [("conv-net", 64, 1e-4), ("conv-blur-net", 32, 1e-4), ("siren", 1024, 1e-3), ("multi-hash", 2048, 3e-4), ("multi-hash", 8192, 3e-4), ("relu-pe", 1024, 1e-4)])
def test_tf32_view_chain_gemms_match_tf32_oracle(env, network, m, tol):
    """conv2d as implicit GEMM (im2col view chain, grouped, replicate padding) and transposed dense GEMMs on the
    gathered tcgen05 kernel: one training step against the oracle with TF32 truncation on exactly those MatMuls."""
    env.set_tf32(True)
    ex = env.example(network, m)
    rng = np.random.default_rng(8)
    params = init_example_params(ex, rng, siren=(network == "siren"))
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    nodes = tensor_core_nodes(env, ex)
    assert nodes, "no MatMul went to the tensor cores"
    upload(env, params)
    from helpers import fill_missing_inputs
    fill_missing_inputs(env, ex.train_graph_json, params)
    env.run(ex.train_graph, 5)
    want = run_graph(ex.train_graph_json, params, 5, tf32=("trunc", nodes))
    worst = {}
    for i, p in enumerate(ex.parameters):
        m_state = ex.optimizer_state[1 + 2 * i]
        worst[p.name() + "#%d" % p.id] = max_rel_err(env.read(m_state), want[m_state.id])
    print(network, worst)
    assert max(worst.values()) <= tol, worst
    assert abs(env.read_parameter_scalar(ex.loss_sum) - float(want[ex.loss_sum.id][0])) <= 1e-5 * abs(float(want[ex.loss_sum.id][0]))


def test_tf32_conv_net_100_steps_within_stated_tolerance(env, built_library):
    """BASELINE.json: "within a stated TF32 tolerance on loss and parameters after N steps for tensor-core GEMMs", stated on
    the headline network.  N = 100 Adam steps of conv-net at m = 256 (fresh synthetic batch and dropout seed every step)
    on the GPU twice -- strict FP32 kernels and TF32 tensor-core operands -- and in the STRICT oracle (oracle.cpu_ref: float32
    sums like the reference's kernels), all fed the same batches and seeds.  The contract:
      * accumulated loss over the 100 steps: strict path 1e-4 relative (measured 4e-6), TF32 path 2e-3 (measured 2.4e-4;
        SURVEY.md section 8d proposal);
      * accumulated accuracy count: 5e-2 relative, both paths (measured 0.6 % / 2.0 %);
      * parameters: Adam normalises each step to ~lr whatever the gradient's size, so an entry whose gradient is a
        cancellation residue moves by +-lr per step in either implementation, and after 100 steps two FP32 implementations
        that differ only in summation order already disagree by up to 0.19 of a tensor's norm (the last layer's 10 biases;
        0.001-0.04 elsewhere) while their losses agree to 4e-6.  A bound on TF32 relative to the oracle alone would
        therefore measure Adam, not TF32: the yardstick is the strict GPU path's own drift on the same run.  Per tensor,
        |theta_tf32 - theta_oracle|_2 / |theta_oracle|_2 <= CONV_NET_TF32_DRIFT_FACTOR * (the strict path's figure) +
        CONV_NET_TF32_DRIFT_FLOOR (measured ratios 1.05 .. 5.3, the largest on a tensor whose strict drift is 0.001), and
        never above CONV_NET_TF32_PARAM_RMS_CAP."""
    from oracle import cpu_ref
    m, steps = 256, 100
    runs = {}
    for mode in ("strict", "tf32"):
        e = built_library.Environment(0) if mode == "tf32" else env
        e.set_tf32(mode == "tf32")
        runs[mode] = (e, e.example("conv-net", m))
    ex = runs["strict"][1]
    rng = np.random.default_rng(77)
    params = init_example_params(ex, rng)
    for e, _ in runs.values():
        upload(e, params)
    program = cpu_ref.Program(ex.train_graph_json)
    state = {pid: np.ascontiguousarray(v, np.float32) for pid, v in params.items()}
    for step in range(steps):
        x, y = synthetic_batch(ex, rng)
        seed = int(rng.integers(0, 2 ** 32))
        for e, exm in runs.values():
            e.write(exm.x, x)
            e.write(exm.y, y)
            e.run(exm.train_graph, seed)
        state[ex.x.id], state[ex.y.id] = x, y
        out, _ = program.run(state, seed)
        state.update({pid: v.copy() for pid, v in out.items()})
    program.close()
    want, acc_want = float(state[ex.loss_sum.id].reshape(-1)[0]), float(state[ex.accuracy_sum.id].reshape(-1)[0])
    rms, loss_err, acc_err = {}, {}, {}
    for mode, (e, exm) in runs.items():
        loss_err[mode] = abs(e.read_parameter_scalar(exm.loss_sum) - want) / abs(want)
        acc_err[mode] = abs(e.read_parameter_scalar(exm.accuracy_sum) - acc_want) / acc_want
        rms[mode] = {p.name() + "#%d" % p.id: float(np.linalg.norm(e.read(p).astype(np.float64) - state[p.id]) / np.linalg.norm(state[p.id].astype(np.float64)))
                     for p in exm.parameters}
        print("conv-net %s vs strict oracle after %d steps: loss rel %.3g, accuracy rel %.3g, |d theta|_2 / |theta|_2 %s" %
              (mode, steps, loss_err[mode], acc_err[mode], {k: "%.3g" % v for k, v in rms[mode].items()}))
    assert loss_err["strict"] <= 1e-4 and loss_err["tf32"] <= TF32_LOSS_TOL, loss_err
    assert max(acc_err.values()) <= CONV_NET_TF32_ACCURACY_TOL, acc_err
    for name, v in rms["tf32"].items():
        assert v <= CONV_NET_TF32_DRIFT_FACTOR * rms["strict"][name] + CONV_NET_TF32_DRIFT_FLOOR, (name, v, rms["strict"][name])
        assert v <= CONV_NET_TF32_PARAM_RMS_CAP, (name, v)
    runs["tf32"][0].close()
