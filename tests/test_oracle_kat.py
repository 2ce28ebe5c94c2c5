"""Pins the CPU oracle (oracle/) against the reference's own known answers: the device tests of
src/lib.rs:26-231, the array_api assert, and the bit-exact vectors derived in SURVEY.md Appendix D
from kernel_common.glsl:205-216 and examples/image_fit/main.rs:163-175."""
import os

import numpy as np
import pytest

from oracle import interp, run_graph
from reference_kats import ALL_CASES, TEST_RAND_SEED, instantiate


@pytest.mark.parametrize("make_case", ALL_CASES, ids=lambda f: f.__name__)
def test_reference_known_answers(host_env, make_case):
    case = make_case()
    scope, ins, outs = instantiate(host_env, case)
    params = {p.id: data for p, (_, _, data) in zip(ins, case.inputs)}
    result = run_graph(scope.export_json(), params, TEST_RAND_SEED)
    for p, (shape, _, expected) in zip(outs, case.outputs):
        np.testing.assert_array_equal(result[p.id].reshape(-1), expected)  # exact small integers


def test_pcg_vectors():
    got = interp.pcg(np.array([0, 1, 2, 3, 4, 0xFFFFFFFF], dtype=np.uint64))
    assert [hex(int(v)) for v in got] == ["0x7bb2fe2", "0xa8beea3c", "0x7a7ecc88", "0x7f0ef6bc", "0x28780864", "0xe62a4902"]


@pytest.mark.parametrize("uid,index,bits", [(0, 0, 0x3EDFFBBC), (0, 1, 0x3F548BF0), (0, 2, 0x3EA07458), (1, 0, 0x3EF8558C), (3, 1567999, 0x3DE3EAF3)])
def test_rand_vectors(uid, index, bits):
    r = interp.rand_from_index(uid, np.array([index]), TEST_RAND_SEED)
    assert int(r.view(np.uint32)[0]) == bits


HASH_VECTORS = [  # pixel (i, j) on a 1024^2 image, grid, expected (ia, ib, ic, id): SURVEY.md Appendix D
    ((0, 0), 2, 9, 3, (0, 1, 3, 2)), ((0, 0), 43, 1936, 44, (0, 1, 44, 45)), ((0, 0), 80, 4096, 1526263, (0, 1, 2551, 2550)),
    ((0, 0), 512, 4096, 1526263, (0, 1, 2551, 2550)), ((511, 512), 2, 9, 3, (3, 2, 6, 7)), ((511, 512), 43, 1936, 44, (905, 906, 989, 990)),
    ((511, 512), 80, 4096, 1526263, (3775, 3760, 2216, 2215)), ((511, 512), 512, 4096, 1526263, (2047, 1536, 8, 503)),
    ((1023, 1023), 2, 9, 3, (2, 1, 7, 4)), ((1023, 1023), 43, 1936, 44, (1810, 1811, 1870, 1871)),
    ((1023, 1023), 80, 4096, 1526263, (886, 873, 3455, 3424)), ((1023, 1023), 512, 4096, 1526263, (1526, 1545, 4095, 3072)),
    ((300, 700), 2, 9, 3, (3, 2, 6, 7)), ((300, 700), 43, 1936, 44, (1264, 1265, 1316, 1317)),
    ((300, 700), 80, 4096, 1526263, (2573, 2562, 1030, 1033)), ((300, 700), 512, 4096, 1526263, (3876, 3877, 2367, 2366)),
]


def hash_indices_graph(env, grid, rows, stride):
    """The index chain of HashGrid::eval (examples/image_fit/main.rs:163-175) written with the Array API."""
    x = env.static_parameter([1, 2], "x")
    out = [env.static_parameter([1], n) for n in "abcd"]
    scope = env.scope()
    xv = scope.parameter_value(x)
    cf = (xv * 0.5 + 0.5) * float(grid)
    c = cf.into_u32()
    c0, c1 = c.lock_axis(-1, 0, False), c.lock_axis(-1, 1, False)
    idx = [((c0 + 0) ^ (c1 * stride + 0)) % rows, ((c0 + 1) ^ (c1 * stride + 0)) % rows,
           ((c0 + 0) ^ (c1 * stride + stride)) % rows, ((c0 + 1) ^ (c1 * stride + stride)) % rows]
    for p, i in zip(out, idx):
        scope.write_parameter_value(p, i.to_f32_bits())
    return scope, x, out


@pytest.mark.parametrize("pixel,grid,rows,stride,expected", HASH_VECTORS)
def test_hash_grid_indices(host_env, pixel, grid, rows, stride, expected):
    scope, x, out = hash_indices_graph(host_env, grid, rows, stride)
    coords = np.array([(pixel[0] + 0.5) * (2.0 / 1024) - 1.0, (pixel[1] + 0.5) * (2.0 / 1024) - 1.0], np.float32)
    result = run_graph(scope.export_json(), {x.id: coords})
    got = tuple(int(result[p.id].view(np.uint32)[0]) for p in out)
    assert got == expected


def test_try_from_reshape_semantics(host_env):
    """src/shape.rs:658-668: inserting/removing unit axes is a view; the op graph must fold such reshapes
    completely (no kernel for the Mov)."""
    a = host_env.static_parameter([8], "a")
    b = host_env.static_parameter([1, 8, 1], "b")
    scope = host_env.scope()
    scope.write_parameter_value(b, scope.parameter_value(a).reshape([1, 1, 8]).reshape([8, 1, 1]).reshape([1, 8, 1]) + 1.0)
    g = scope.build_graph().export_json()
    assert len(g["clusters"]) == 1 and g["clusters"][0]["label"].startswith("PerElement")
    with pytest.raises(Exception):
        scope.parameter_value(a).reshape([1, 9, 1])


def test_unpad_and_windows_adjoints():
    """Adjoint identities the reference's backward rules rely on (array.rs:954-975): <pad(x), y> == <x, unpad(y)>
    and <windows(x), w> == <x, windows_to_image(w)> for overlapping 3x3/1 windows."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 5, 6, 3)).astype(np.float32)
    pad = 2
    xp = np.pad(x, ((0, 0), (pad, pad), (pad, pad), (0, 0)), mode="edge")
    y = rng.standard_normal(xp.shape).astype(np.float32)
    u = interp._unpad(y.reshape(-1), list(y.shape), 1, pad)
    u = interp._unpad(u, [2, 5, 6 + 2 * pad, 3], 2, pad).reshape(x.shape)
    assert abs(float((xp.astype(np.float64) * y).sum()) - float((x.astype(np.float64) * u).sum())) < 1e-3
    oh, ow = 3, 4
    win = np.stack([np.stack([x[:, fy:fy + oh, fx:fx + ow, :] for fx in range(3)], axis=3) for fy in range(3)], axis=3)  # [m,oh,ow,fh,fw,c]
    win = win[:, :, :, None]  # groups = 1
    w = rng.standard_normal(win.shape).astype(np.float32)
    img = interp._windows_to_image(w.reshape(-1), list(w.shape), list(x.shape), 1, 1).reshape(x.shape)
    assert abs(float((win.astype(np.float64) * w).sum()) - float((x.astype(np.float64) * img).sum())) < 1e-3


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_STEPS = [("linear", 8), ("single-layer-dropout", 8), ("conv-net", 4), ("conv-blur-net", 2)]


def load_golden_step(network, m):
    z = np.load(os.path.join(GOLDEN_DIR, "step_%s_m%d.npz" % (network, m)))
    inputs = {int(k[3:]): z[k] for k in z.files if k.startswith("in_")}
    outputs = {int(k[4:]): z[k] for k in z.files if k.startswith("out_")}
    return inputs, outputs, int(z["seed"][0])


@pytest.mark.parametrize("network,m", GOLDEN_STEPS, ids=[g[0] for g in GOLDEN_STEPS])
def test_oracle_reproduces_golden_steps(host_env, network, m):
    """tests/golden/step_*.npz (made by tests/golden/make_golden.py) pin the oracle's arithmetic: one SGD step of each
    fashion_mnist network, every output.  1e-6 relative leaves room for a different BLAS in the float64 products."""
    inputs, outputs, seed = load_golden_step(network, m)
    ex = host_env.example(network, m, optimizer="descent")
    got = run_graph(ex.train_graph_json, inputs, seed)
    assert set(got) == set(outputs)
    for pid, want in outputs.items():
        scale = max(float(np.abs(want).max()), 1e-30)
        assert float(np.abs(got[pid].astype(np.float64) - want).max()) <= 1e-6 * scale, (network, pid)


def test_integer_vectors_are_bit_exact():
    """pcg hash and Rand (kernel_common.glsl:205-216) against the committed vectors."""
    z = np.load(os.path.join(GOLDEN_DIR, "integer_vectors.npz"))
    np.testing.assert_array_equal(interp.pcg(z["index"]), z["pcg"])
    np.testing.assert_array_equal(interp.rand_from_index(3, z["index"], 7).view(np.uint32), z["rand_uid3_seed7"])


# ---- oracle/cpu_ref.cpp (the C++ port timed as the CPU baseline) against the numpy oracle ---------------------------

@pytest.mark.parametrize("network,m,optimizer", [("conv-net", 4, "descent"), ("conv-blur-net", 2, "descent"), ("single-layer-dropout", 8, "descent"),
                                                 ("multi-hash", 64, "adam")])
@pytest.mark.parametrize("optimised", [False, True], ids=["raw-graph", "after-passes"])
def test_cpu_ref_matches_numpy_oracle(host_env, network, m, optimizer, optimised):
    """Every output of one training step, raw graph and the graph after the frontend's passes (what bench.py times).
    1e-4 of each tensor's maximum: cpu_ref sums sequentially in f32 like the reference's kernels, the oracle in f64.
    (SGD for the fashion_mnist nets: Adam's first step is ill-conditioned in near-zero gradient entries, see
    test_gpu_networks.py.)"""
    from helpers import init_example_params, synthetic_batch
    from oracle import cpu_ref
    ex = host_env.example(network, m, optimizer=optimizer)
    rng = np.random.default_rng(11)
    params = init_example_params(ex, rng)
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    for node in ex.train_graph_json["nodes"]:
        if node["op"] == "Input" and node["parameter"] not in params:
            shape = host_env.parameter(node["parameter"]).shape()
            params[node["parameter"]] = np.full(shape, 1.0 / 16.0, np.float32)
    want = run_graph(ex.train_graph_json, params, 3)
    got, seconds = cpu_ref.run_graph(ex.train_graph.export_json() if optimised else ex.train_graph_json, params, 3, threads=4)
    assert set(got) == set(want) and seconds >= 0
    for pid, w in want.items():
        scale = max(float(np.abs(w).max()), 1e-30)
        assert float(np.abs(got[pid].astype(np.float64) - w).max()) <= 1e-4 * scale, (network, pid)


@pytest.mark.parametrize("make_case", ALL_CASES, ids=lambda f: f.__name__)
def test_cpu_ref_reference_known_answers(host_env, make_case):
    """The reference's own device tests (src/lib.rs:26-231) through the C++ port, exact."""
    from oracle import cpu_ref
    case = make_case()
    scope, ins, outs = instantiate(host_env, case)
    got, _ = cpu_ref.run_graph(scope.export_json(), {p.id: data for p, (_, _, data) in zip(ins, case.inputs)}, TEST_RAND_SEED, threads=2)
    for p, (_, _, expected) in zip(outs, case.outputs):
        np.testing.assert_array_equal(got[p.id].reshape(-1), np.asarray(expected, np.float32).reshape(-1))


@pytest.mark.parametrize("network,m,tf32", [("conv-net", 16, False), ("conv-net", 16, True), ("single-layer-dropout", 32, True)])
def test_cpu_ref_checker_mode_matches_numpy_oracle(host_env, network, m, tf32):
    """cpu_ref's checker mode (float64-accumulated sums, TF32 truncation on chosen MatMuls) is the numpy interpreter's
    arithmetic on all host cores: the large-batch GPU parity tests (m = 1000 / 8192, tests/test_gpu_bench_regime.py) rely
    on it where the interpreter would take minutes.  Every output of one SGD step within 2e-6 of each tensor's maximum
    (both round a float64 sum once; only the order of the float64 additions differs).  With TF32 truncation 2e-5: a last-bit
    difference in one layer's output can cross a truncation boundary of the next layer's operand (2^-10 of that element)."""
    from helpers import init_example_params, synthetic_batch
    from oracle import cpu_ref
    ex = host_env.example(network, m, optimizer="descent")
    rng = np.random.default_rng(12)
    params = init_example_params(ex, rng)
    params[ex.x.id], params[ex.y.id] = synthetic_batch(ex, rng)
    matmuls = {n["id"] for n in ex.train_graph_json["nodes"] if n["op"] == "MatMul"}
    nodes = set(sorted(matmuls)[::2]) if tf32 else set()  # every other MatMul: the per-node selection is exercised too
    want = run_graph(ex.train_graph_json, params, 3, tf32=("trunc", nodes) if nodes else None)
    got = cpu_ref.check_graph(ex.train_graph_json, params, 3, tf32_nodes=nodes, threads=4)
    assert set(got) == set(want)
    for pid, w in want.items():
        scale = max(float(np.abs(w).max()), 1e-30)
        assert float(np.abs(got[pid].astype(np.float64) - w).max()) <= (2e-5 if tf32 else 2e-6) * scale, (network, pid)
