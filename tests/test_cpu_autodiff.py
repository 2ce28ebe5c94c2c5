"""Independent check of the hand-written backward rules (src/array.rs:794-1214 restated in
descent_b200/csrc/array.cpp): the oracle evaluates the exported training graph with plain SGD (lr = 1,
no momentum, no weight decay), so new_theta = theta - dL/dtheta; torch float64 autograd computes the same
gradient from an independent forward model (replicate padding, grouped conv, max-pool ties, dropout mask from
the restated pcg hash)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import interp, run_graph


def sgd_graph(env, build_loss, params):
    """theta <- theta - 1.0 * grad for every trainable parameter the loss touches."""
    import descent_b200 as d
    scope = env.scope()
    build_loss(scope)
    ps = scope.trainable_parameters()
    d.StochasticGradientDescent(env, scope, ps, 1.0, 0.0)
    return scope.export_json(), ps


def grads_from_oracle(graph, values, ps, seed=0):
    out = run_graph(graph, values, seed)
    return {p.id: values[p.id].astype(np.float64) - out[p.id].astype(np.float64) for p in ps}


def t64(a):
    return torch.tensor(np.asarray(a, np.float64), requires_grad=True)


def test_dense_leaky_softmax_xent(host_env):
    import descent_b200 as d
    env = host_env
    m, i, h, o = 12, 7, 5, 4
    x = env.static_parameter([m, i], "x")
    y = env.static_parameter([m, 1], "y")
    fc1, fc2 = d.Dense(env, i, h), d.Dense(env, h, o)

    def loss(scope):
        z = fc2.train(fc1.train(scope.parameter(x)).leaky_relu(0.01))
        d.softmax_cross_entropy_loss(z, y).set_loss()
    graph, ps = sgd_graph(env, loss, None)
    rng = np.random.default_rng(0)
    vals = {p.id: rng.standard_normal(p.shape()).astype(np.float32) for p in ps}
    vals[x.id] = rng.standard_normal((m, i)).astype(np.float32)
    vals[y.id] = rng.integers(0, o, (m, 1)).astype(np.float32)
    got = grads_from_oracle(graph, vals, ps)
    w1, b1, w2, b2 = (t64(vals[p.id]) for p in ps)
    z = F.leaky_relu(t64(vals[x.id]).detach() @ w1 + b1, 0.01) @ w2 + b2
    F.cross_entropy(z, torch.tensor(vals[y.id][:, 0].astype(np.int64)), reduction="mean").backward()
    for p, t in zip(ps, (w1, b1, w2, b2)):
        np.testing.assert_allclose(got[p.id], t.grad.numpy(), rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize("groups,stride,pad", [(1, 1, 1), (2, 1, 1), (1, 2, 0), (4, 2, 1)])
def test_conv2d_replicate_pad_grouped(host_env, groups, stride, pad):
    """Conv2D with clamp-to-edge padding and the corrected col2im adjoint (A.9)."""
    import descent_b200 as d
    env = host_env
    m, hh, ww, ic, oc, f = 3, 9 if stride == 2 and pad == 0 else 8, 9 if stride == 2 and pad == 0 else 8, 4, 8, 3
    if stride == 2 and pad == 1:
        hh = ww = 7  # (7 + 2 - 3) / 2 + 1 = 4 exactly
    x = env.trainable_parameter([m, hh, ww, ic], "x")
    conv = d.Conv2D(env, ic, oc, f, f, pad=pad, stride=(stride, stride), groups=groups)

    def loss(scope):
        yv = conv.train(scope.parameter(x)).leaky_relu(0.1)
        (yv * yv).reduce_sum(-1, True).reduce_sum(-2, True).reduce_sum(-3, True).reshape([m, 1]).set_loss()
    graph, ps = sgd_graph(env, loss, None)
    rng = np.random.default_rng(groups * 10 + stride)
    vals = {p.id: rng.standard_normal(p.shape()).astype(np.float32) for p in ps}
    got = grads_from_oracle(graph, vals, ps)
    by_name = {p.name(): p for p in ps}
    xt, ft, bt = t64(vals[by_name["x"].id]), t64(vals[by_name["f"].id]), t64(vals[by_name["b"].id])
    xin = xt.permute(0, 3, 1, 2)
    if pad:
        xin = F.pad(xin, (pad, pad, pad, pad), mode="replicate")
    g, ocg, fh, fw, icg = ft.shape
    weight = ft.permute(0, 1, 4, 2, 3).reshape(g * ocg, icg, fh, fw)  # filter [g, oc/g, fh, fw, ic/g] -> OIHW
    out = F.conv2d(xin, weight, bias=bt, stride=stride, groups=groups)
    out = F.leaky_relu(out, 0.1)
    ((out * out).sum() / m).backward()
    for name, t in (("x", xt), ("f", ft), ("b", bt)):
        np.testing.assert_allclose(got[by_name[name].id], t.grad.numpy(), rtol=3e-4, atol=3e-5, err_msg=name)


def test_max_pool_dropout(host_env):
    import descent_b200 as d
    env = host_env
    m, hh, ww, c = 2, 6, 6, 3
    x = env.trainable_parameter([m, hh, ww, c], "x")
    pool, drop = d.MaxPool2D(env), d.Dropout(env, 0.5)
    seed = 1234

    def loss(scope):
        yv = drop.train(pool.train(scope.parameter(x)).flatten())
        (yv * yv).reduce_sum(-1, True).set_loss()
    graph, ps = sgd_graph(env, loss, None)
    rng = np.random.default_rng(5)
    xv = rng.permutation(m * hh * ww * c).astype(np.float32).reshape(m, hh, ww, c) / 7.0  # distinct values: no pooling ties
    got = grads_from_oracle(graph, {x.id: xv}, ps, seed)
    xt = t64(xv)
    pooled = F.max_pool2d(xt.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(m, -1)
    r = interp.rand_from_index(0, np.arange(pooled.numel()), seed).reshape(pooled.shape)
    mask = torch.tensor((r > np.float32(0.5)).astype(np.float64))
    yv = pooled * mask * 2.0
    ((yv * yv).sum() / m).backward()
    np.testing.assert_allclose(got[x.id], xt.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_siren_positional_encoding_mse(host_env):
    ex = host_env.example("relu-pe", 32)
    import descent_b200 as d
    env = host_env
    scope = env.scope()
    # same network, plain SGD(1.0): rebuild the loss of examples/image_fit/main.rs:308-313 on the example's module graph
    graph = ex.train_graph_json  # Adam graph: use m-state to recover gradients instead (m = 0.1 * g after one step)
    rng = np.random.default_rng(9)
    from helpers import init_example_params, synthetic_batch
    params = init_example_params(ex, rng)
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    out = run_graph(graph, params, 0)
    ws = [t64(params[p.id]) for p in ex.parameters]
    x = torch.tensor(params[ex.x.id].astype(np.float64))
    freq = (2.0 ** torch.arange(8, dtype=torch.float64)) * np.float64(np.float32(np.pi))
    phase = torch.arange(2, dtype=torch.float64).reshape(2, 1) * 0.5 * np.float64(np.float32(np.pi))
    h = torch.sin(x.reshape(-1, 2, 1, 1) * freq + phase).reshape(-1, 32)
    for li in range(4):
        h = F.leaky_relu(h @ ws[2 * li] + ws[2 * li + 1], 0.01)
    pred = h @ ws[8] + ws[9]
    (((pred - torch.tensor(params[ex.y.id].astype(np.float64))) ** 2).sum() / x.shape[0]).backward()
    for i, p in enumerate(ex.parameters):
        m_state = out[ex.optimizer_state[1 + 2 * i].id].astype(np.float64) / np.float64(np.float32(1.0) - np.float32(0.9))
        np.testing.assert_allclose(m_state, ws[i].grad.numpy(), rtol=2e-3, atol=1e-6 * max(1.0, float(ws[i].grad.abs().max())))


@pytest.mark.parametrize("network,m,kwargs", [("linear", 16, {}), ("single-layer-dropout", 16, {}), ("conv-net", 6, {}), ("conv-blur-net", 3, {}), ("relu-pe", 32, {}),
                                              ("siren", 32, {}), ("multi-hash", 48, {}), ("sentiment", 4, {"image_width": 40, "image_height": 5})])
def test_graph_passes_preserve_every_output_bit_for_bit(host_env, network, m, kwargs):
    """The graph the backend compiles (after dead-code / move elimination, x*1 and x+0 simplification, CSE, all-reduce
    view hoisting, permutation sinking and activation-sign reuse, descent_b200/csrc/graph.cpp) must compute exactly
    what the raw graph computes: the oracle interprets both and every output has to be identical, bit for bit."""
    from helpers import init_example_params, synthetic_batch
    ex = host_env.example(network, m, **kwargs)
    rng = np.random.default_rng(len(network) + m)
    params = init_example_params(ex, rng, siren=(network == "siren"))
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    for node in ex.train_graph_json["nodes"]:
        if node["op"] == "Input" and node["parameter"] not in params:
            params[node["parameter"]] = np.full(host_env.parameter(node["parameter"]).shape(), 1.0 / 16.0, np.float32)
    raw = run_graph(ex.train_graph_json, params, 9)
    optimised = run_graph(ex.train_graph.export_json(), params, 9)
    assert set(raw) == set(optimised)
    for pid, want in raw.items():
        np.testing.assert_array_equal(optimised[pid].view(np.uint32), want.view(np.uint32), err_msg="%s parameter %d" % (network, pid))
