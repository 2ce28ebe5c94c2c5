"""The reference's device tests on the CUDA backend (src/lib.rs:26-231): bit-exact, through the C ABI."""
import numpy as np
import pytest

import descent_b200 as d
from reference_kats import ALL_CASES, TEST_RAND_SEED, instantiate
from test_oracle_kat import HASH_VECTORS, hash_indices_graph

pytestmark = pytest.mark.gpu


def test_parameters(env):  # src/lib.rs:26-34
    data = np.arange(10, dtype=np.float32)
    a = env.static_parameter_with_data([10], "a", data)
    np.testing.assert_array_equal(env.read_parameter_to_vec(a), data)


def test_writer_zero_fills_tail(env):  # staging.rs:181-187
    a = env.static_parameter_with_data([16], "a", np.ones(16, np.float32))
    env.write(a, np.full(5, 7.0, np.float32))
    np.testing.assert_array_equal(env.read_parameter_to_vec(a), np.r_[np.full(5, 7.0), np.zeros(11)].astype(np.float32))
    env.write(a, np.ones(16, np.float32))
    env.zero_fill(a)
    np.testing.assert_array_equal(env.read_parameter_to_vec(a), np.zeros(16, np.float32))


def test_prefetch_lands_at_the_next_run(env):
    """Environment.prefetch_pinned: batch i+1 crosses PCIe on the copy stream while batch i is in use; the parameter
    only changes at the next run (the staging ring of staging.rs:100-140 gives the reference the same overlap)."""
    n = 1 << 20
    x = env.static_parameter([n], "x")
    y = env.static_parameter([n], "y")
    scope = env.scope()
    scope.write_parameter_value(y, scope.parameter_value(x) * 2.0 + 1.0)
    g = scope.build_graph()
    host = d.pinned_array(n)
    first = np.arange(n, dtype=np.float32)
    env.write(x, first)
    for step in range(4):
        host[:] = np.float32(step) - first
        env.prefetch_pinned(x, host)
        if step == 0:  # nothing ran yet: reading the parameter itself forces the pending copy to land
            np.testing.assert_array_equal(env.read_parameter_to_vec(x), host)
        env.run(g, TEST_RAND_SEED)
        np.testing.assert_array_equal(env.read_parameter_to_vec(y), host * np.float32(2.0) + np.float32(1.0))
    with pytest.raises(d.DescentError):
        env.prefetch_pinned(x, d.pinned_array(n // 2))  # whole parameters only


@pytest.mark.parametrize("use_cuda_graph", [True, False])
@pytest.mark.parametrize("make_case", ALL_CASES, ids=lambda f: f.__name__)
def test_reference_known_answers(env, make_case, use_cuda_graph):
    env.set_options(use_cuda_graph=use_cuda_graph)
    case = make_case()
    scope, ins, outs = instantiate(env, case)
    for p, (_, _, data) in zip(ins, case.inputs):
        env.write(p, data)
    g = scope.build_graph()
    for _ in range(2):  # second run replays the captured step
        env.run(g, TEST_RAND_SEED)
    for p, (_, _, expected) in zip(outs, case.outputs):
        np.testing.assert_array_equal(env.read_parameter_to_vec(p), expected)


@pytest.mark.parametrize("pixel,grid,rows,stride,expected", HASH_VECTORS[::3])
def test_hash_grid_indices(env, pixel, grid, rows, stride, expected):
    scope, x, out = hash_indices_graph(env, grid, rows, stride)
    env.write(x, np.array([(pixel[0] + 0.5) * (2.0 / 1024) - 1.0, (pixel[1] + 0.5) * (2.0 / 1024) - 1.0], np.float32))
    env.run(scope.build_graph(), 0)
    got = tuple(int(env.read_parameter_to_vec(p).view(np.uint32)[0]) for p in out)
    assert got == expected


def test_rand_is_bit_exact(env):
    """Rand{uid} = float(pcg(pcg(index)+seed+uid)) * 2^-32 (kernel_common.glsl:205-216)."""
    from oracle import interp
    out = env.static_parameter([1000, 7], "r")
    scope = env.scope()
    scope.rand([3]).value()  # uid 0 is consumed elsewhere; the tested node gets uid 1
    scope.write_parameter_value(out, scope.rand([1000, 7]).value() * 1.0)
    env.run(scope.build_graph(), TEST_RAND_SEED)
    want = interp.rand_from_index(1, np.arange(7000), TEST_RAND_SEED)
    np.testing.assert_array_equal(env.read_parameter_to_vec(out).view(np.uint32), want.view(np.uint32))
