"""Conv2D forward, backward-input and weight-gradient kernels on their own (halo-tiled tcgen05 kernels with TF32
on, gathered / streaming strict-FP32 kernels with TF32 off), against the oracle's evaluation of the same graph.

One SGD step with lr = 1 turns every parameter into theta - dL/dtheta, so comparing parameters compares
gradients; the loss is quadratic in the convolution's output, so the forward result enters every gradient.  Reference semantics: array.rs conv2d / image_to_windows,
kernel.rs MatMulKernel, WindowsToImageKernel (712-810), UnpadKernel (644-710)."""
import numpy as np
import pytest

import descent_b200 as d
from helpers import max_rel_err
from oracle import run_graph

pytestmark = pytest.mark.gpu

# TF32 operands carry 10 explicit mantissa bits (hardware truncates, pinned by test_gpu_gemm_tf32.py): relative to the
# largest gradient entry, sums of a few thousand products stay within 2e-3; strict FP32 within 1e-5 (north_star)
TF32_TOL = 2e-3
FP32_TOL = 1e-5

SHAPES = [
    # images, height, width, in channels, out channels, groups
    (8, 14, 14, 16, 32, 2),   # conv-net's second convolution: padded width 16 -> halo kernels
    (4, 14, 14, 8, 16, 1),    # single group
    (3, 14, 14, 32, 32, 2),   # 96 weight-gradient rows: the M=128 form of that kernel
    (2, 6, 6, 8, 8, 1),       # padded width 8: 16 image rows per tile, images smaller than a tile
    (3, 30, 30, 8, 16, 2),    # padded width 32, ragged last tile
    (5, 12, 12, 8, 16, 1),    # padded width 14 does not divide 128 -> gathered kernels
    (6, 28, 28, 1, 16, 1),    # one input channel: streaming FP32 kernels
]


@pytest.mark.parametrize("tf32", [True, False], ids=["tf32", "strict"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv2d_step_matches_oracle(env, shape, tf32):
    m, hh, ww, ic, oc, groups = shape
    env.set_tf32(tf32)
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

    rng = np.random.default_rng(hash(shape) % 1000)
    vals = {p.id: rng.standard_normal(p.shape()).astype(np.float32) for p in ps}
    vals[w_out.id] = rng.standard_normal(w_out.shape()).astype(np.float32)
    for pid, v in vals.items():
        env.write(env.parameter(pid), v)
    env.run(graph, 1)
    want = run_graph(graph_json, vals, 1)
    tol = TF32_TOL if tf32 else FP32_TOL
    errs = {}
    for p in ps:
        grad_got = vals[p.id].astype(np.float64) - env.read(p).astype(np.float64)
        grad_want = vals[p.id].astype(np.float64) - want[p.id].astype(np.float64)
        errs[p.name()] = max_rel_err(grad_got, grad_want)
        if errs[p.name()] > tol:
            bad = np.argwhere(np.abs(grad_got - grad_want) > tol * np.abs(grad_want).max())
            print(p.name(), "mismatches", len(bad), "of", grad_want.size, "first", bad[:6].tolist(),
                  "got", [float(grad_got[tuple(b)]) for b in bad[:6]], "want", [float(grad_want[tuple(b)]) for b in bad[:6]])
    print(shape, "tf32" if tf32 else "strict", errs)
    assert max(errs.values()) <= tol, errs
