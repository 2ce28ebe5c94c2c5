"""Op-level entry points of the C ABI (include/descent_api.h dsc_op_*): conv2d forward / backward,
deterministic scatter_add, softmax cross-entropy and the multi-tensor Adam step, each against a direct numpy statement of
the reference's op (array.rs:989-1031, kernel.rs:812-874, loss.rs:4-34, optimizer.rs:62-112)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def conv2d_reference(x, f, pad, stride):
    """NHWC, replicate padding, filter [groups, oc_per_group, fh, fw, ic_per_group] (array.rs:989-1031), float64."""
    m, h, w, ic = x.shape
    g, ocg, fh, fw, icg = f.shape
    xp = np.pad(x.astype(np.float64), ((0, 0), (pad, pad), (pad, pad), (0, 0)), mode="edge")
    oh, ow = (h + 2 * pad - fh) // stride[1] + 1, (w + 2 * pad - fw) // stride[0] + 1
    y = np.zeros((m, oh, ow, g * ocg))
    for gi in range(g):
        for fy in range(fh):
            for fx in range(fw):
                patch = xp[:, fy:fy + (oh - 1) * stride[1] + 1:stride[1], fx:fx + (ow - 1) * stride[0] + 1:stride[0], gi * icg:(gi + 1) * icg]
                y[..., gi * ocg:(gi + 1) * ocg] += patch @ f[gi, :, fy, fx, :].astype(np.float64).T
    return y


@pytest.mark.parametrize("tf32", [False, True], ids=["strict", "tf32"])
@pytest.mark.parametrize("shape", [(64, 14, 14, 16, 32, 3, 3, 1, 2), (40, 28, 28, 1, 16, 3, 3, 1, 1), (8, 9, 11, 4, 6, 3, 3, 0, 1)], ids=str)
def test_op_conv2d_forward_and_backward(env, shape, tf32):
    m, h, w, ic, oc, fh, fw, pad, groups = shape
    env.set_tf32(tf32)
    rng = np.random.default_rng(11)
    x = rng.standard_normal((m, h, w, ic)).astype(np.float32)
    f = (rng.standard_normal((groups, oc // groups, fh, fw, ic // groups)) * 0.2).astype(np.float32)
    if tf32:  # the tensor cores truncate operands to 10 mantissa bits: give them operands that survive unchanged
        x = (x.view(np.uint32) & 0xFFFFE000).view(np.float32)
        f = (f.view(np.uint32) & 0xFFFFE000).view(np.float32)
    fwd = env.op_conv2d(m, h, w, ic, oc, fh, fw, pad=pad, groups=groups)
    fwd.write("x", x)
    fwd.write("filter", f)
    fwd.run()
    want = conv2d_reference(x, f, pad, (1, 1))
    got = fwd.read("y")
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    address, size = fwd.device_buffer("y")
    assert address != 0 and size == want.size * 4
    # backward: dx and dfilter as the adjoints of the forward map (checked through inner products with random probes)
    bwd = env.op_conv2d(m, h, w, ic, oc, fh, fw, pad=pad, groups=groups, backward=True)
    dy = rng.standard_normal(want.shape).astype(np.float32)
    if tf32:
        dy = (dy.view(np.uint32) & 0xFFFFE000).view(np.float32)
    bwd.write("x", x)
    bwd.write("filter", f)
    bwd.write("dy", dy)
    bwd.run()
    dx, df = bwd.read("dx").astype(np.float64), bwd.read("dfilter").astype(np.float64)
    px, pf = rng.standard_normal(x.shape), rng.standard_normal(f.shape)
    lhs_x = (conv2d_reference(px, f, pad, (1, 1)) * dy).sum()  # <conv(px, f), dy> = <px, dx>
    lhs_f = (conv2d_reference(x, pf, pad, (1, 1)) * dy).sum()  # <conv(x, pf), dy> = <pf, dfilter>
    tol = 2e-3 if tf32 else 1e-4  # tf32: the probes' products are exact only on the truncated operands' side
    assert abs((px * dx).sum() - lhs_x) <= tol * max(1.0, abs(lhs_x))
    assert abs((pf * df).sum() - lhs_f) <= tol * max(1.0, abs(lhs_f))


def test_op_scatter_add_is_exact_and_deterministic(env):
    rows, inner, count = 577, 2, 100000
    rng = np.random.default_rng(4)
    table = rng.standard_normal((rows, inner)).astype(np.float32)
    values = rng.standard_normal((count, inner)).astype(np.float32)
    indices = rng.integers(0, rows, count).astype(np.uint32)
    op = env.op_scatter_add(rows, inner, count)
    results = []
    for _ in range(2):
        op.write("table", table)
        op.write("values", values)
        op.write("indices", indices.view(np.float32))
        op.run()
        results.append(op.read("table"))
    np.testing.assert_array_equal(results[0].view(np.uint32), results[1].view(np.uint32))  # bitwise reproducible
    want = table.astype(np.float64)
    np.add.at(want, indices, values.astype(np.float64))
    assert np.abs(results[0] - want).max() <= 1e-5 * np.abs(want).max()


def test_op_softmax_cross_entropy(env):
    rows, classes = 1000, 10
    rng = np.random.default_rng(9)
    z = (rng.standard_normal((rows, classes)) * 3).astype(np.float32)
    y = rng.integers(0, classes, (rows, 1)).astype(np.float32)
    op = env.op_softmax_cross_entropy(rows, classes)
    op.write("z", z)
    op.write("y", y)
    op.run()
    z64 = z.astype(np.float64)
    p = np.exp(z64 - z64.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    picked = p[np.arange(rows), y[:, 0].astype(int)]
    np.testing.assert_allclose(op.read("loss")[:, 0], -np.log(picked), rtol=2e-5, atol=1e-6)
    np.testing.assert_array_equal(op.read("accuracy")[:, 0], (z.argmax(-1) == y[:, 0]).astype(np.float32))
    onehot = np.eye(classes)[y[:, 0].astype(int)]
    np.testing.assert_allclose(op.read("dz"), (p - onehot) / rows, atol=2e-9, rtol=2e-5)  # set_loss: gradient of the batch MEAN (array.rs set_loss)


def test_op_adam_step_multi_tensor(env):
    counts = [1000, 37, 4096]
    lr, b1, b2, eps = 0.01, 0.9, 0.999, 1e-8
    rng = np.random.default_rng(2)
    op = env.op_adam_step(counts, lr, b1, b2, eps)
    theta = [rng.standard_normal(n).astype(np.float32) for n in counts]
    m = [np.zeros(n) for n in counts]
    v = [np.zeros(n) for n in counts]
    want = [t.astype(np.float64) for t in theta]
    for i, t in enumerate(theta):
        op.write("theta%d" % i, t)
    for step in range(1, 4):
        grads = [rng.standard_normal(n).astype(np.float32) for n in counts]
        for i, g in enumerate(grads):
            op.write("grad%d" % i, g)
        op.run()
        alpha = lr * np.sqrt(1 - b2 ** step) / (1 - b1 ** step)  # optimizer.rs:96-103
        for i, g in enumerate(grads):
            m[i] = b1 * m[i] + (1 - b1) * g
            v[i] = b2 * v[i] + (1 - b2) * g.astype(np.float64) ** 2
            want[i] = want[i] - alpha * m[i] / (np.sqrt(v[i]) + eps)
    for i in range(len(counts)):
        np.testing.assert_allclose(op.read("theta%d" % i), want[i], rtol=2e-5, atol=2e-6)
