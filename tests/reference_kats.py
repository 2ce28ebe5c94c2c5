"""The reference's own device tests (src/lib.rs:26-231) and the array_api assert
(examples/array_api/main.rs:24), restated against the Python mirror of the API.  Each case returns
(graph builder, input data, expected outputs) so the same definition pins the CPU oracle
(tests/test_oracle_kat.py) and the CUDA backend (tests/test_gpu_kat.py)."""
import numpy as np

TEST_RAND_SEED = 0x5EED5EED  # src/lib.rs:24


class Case:
    def __init__(self, name):
        self.name = name
        self.inputs = []   # (shape, name, data)
        self.outputs = []  # (shape, name, expected)
        self.build = None


def _case(name):
    return Case(name)


def reduce_case():  # src/lib.rs:36-55
    c = _case("reduce")
    a = np.arange(100, dtype=np.float32)
    c.inputs = [([10, 10], "a", a)]
    c.outputs = [([10, 1], "b", a.reshape(10, 10).sum(axis=1))]
    c.build = lambda scope, ins, outs: scope.write_parameter_value(outs[0], scope.parameter_value(ins[0]).reduce_sum(-1, True))
    return c


def pad_image_case():  # src/lib.rs:57-73
    c = _case("pad_image")
    c.inputs = [([1, 8, 8, 1], "a", np.ones(64, np.float32))]
    c.outputs = [([1, 10, 10, 1], "b", np.ones(100, np.float32))]
    c.build = lambda scope, ins, outs: scope.write_parameter_value(outs[0], scope.parameter_value(ins[0]).pad_image(1))
    return c


def unpad_image_case():  # src/lib.rs:75-98
    c = _case("unpad_image")
    unpad = lambda a: 2.0 if a in (0, 7) else 1.0
    expected = np.array([unpad(y) * unpad(x) for y in range(8) for x in range(8)], np.float32)
    c.inputs = [([1, 10, 10, 1], "a", np.ones(100, np.float32))]
    c.outputs = [([1, 8, 8, 1], "b", expected)]
    c.build = lambda scope, ins, outs: scope.write_parameter_value(outs[0], scope.parameter_value(ins[0]).unpad_image(1))
    return c


def conv2d_case():  # src/lib.rs:100-124
    c = _case("conv2d")
    c.inputs = [([1, 10, 10, 1], "a", np.ones(100, np.float32)), ([1, 1, 3, 3, 1], "b", np.ones(9, np.float32))]
    c.outputs = [([1, 8, 8, 1], "c", np.full(64, 9.0, np.float32))]
    c.build = lambda scope, ins, outs: scope.write_parameter_value(outs[0], scope.parameter(ins[0]).conv2d(ins[1], 0, (1, 1)).value())
    return c


def max_pool2d_case():  # src/lib.rs:126-147
    c = _case("max_pool2d")
    expected = np.array([11 + 2 * (i % 5) + 20 * (i // 5) for i in range(25)], np.float32)
    c.inputs = [([1, 10, 10, 1], "a", np.arange(100, dtype=np.float32))]
    c.outputs = [([1, 5, 5, 1], "b", expected)]
    c.build = lambda scope, ins, outs: scope.write_parameter_value(outs[0], scope.parameter(ins[0]).max_pool2d((2, 2), (2, 2)).value())
    return c


def gather_case():  # src/lib.rs:149-173
    c = _case("gather")
    c.inputs = [([1, 200, 1], "a", np.array([i * i for i in range(200)], np.float32)),
                ([100], "b", np.array([99 - i for i in range(100)], np.float32))]
    c.outputs = [([1, 100, 1], "c", np.array([(99 - i) * (99 - i) + 1 for i in range(100)], np.float32))]
    c.build = lambda scope, ins, outs: scope.write_parameter_value(
        outs[0], scope.parameter_value(ins[0]).gather(1, scope.parameter_value(ins[1]).into_u32()) + 1.0)
    return c


def scatter_add_case():  # src/lib.rs:175-202
    c = _case("scatter_add")
    c.inputs = [([1, 100, 1], "a", np.ones(100, np.float32)), ([100], "b", np.array([i % 10 for i in range(100)], np.float32))]
    c.outputs = [([1, 10, 1], "c", np.full(10, 10.0, np.float32))]
    c.build = lambda scope, ins, outs: scope.write_parameter_value(
        outs[0], scope.literal(0.0).value().broadcast([1, 10, 1]).scatter_add(ins[0], -2, scope.parameter_value(ins[1]).into_u32()))
    return c


def concat_case():  # src/lib.rs:204-231
    c = _case("concat")
    a = np.array([i for i in range(200) if ((i // 10) & 1) == 0], np.float32)
    b = np.array([i for i in range(200) if ((i // 10) & 1) == 1], np.float32)
    c.inputs = [([10, 10], "a", a), ([10, 10], "b", b)]
    c.outputs = [([10, 20], "c", np.arange(200, dtype=np.float32))]
    c.build = lambda scope, ins, outs: scope.write_parameter_value(outs[0], scope.parameter_value(ins[0]).concat(ins[1], -1))
    return c


def array_api_case():  # examples/array_api/main.rs:10-24: 2*(I.x) + y*y + 1 = [10, 15, 22]
    c = _case("array_api")
    c.inputs = [([3, 3], "m", np.eye(3, dtype=np.float32)), ([3, 1], "x", np.array([4, 5, 6], np.float32)),
                ([3, 1], "y", np.array([1, 2, 3], np.float32))]
    c.outputs = [([3, 1], "z", np.array([10, 15, 22], np.float32))]

    def build(scope, ins, outs):
        m, x, y = (scope.parameter_value(p) for p in ins)
        scope.write_parameter_value(outs[0], 2.0 * m.matmul(x) + y * y + 1.0)
    c.build = build
    return c


ALL_CASES = [reduce_case, pad_image_case, unpad_image_case, conv2d_case, max_pool2d_case, gather_case, scatter_add_case, concat_case,
             array_api_case]


def instantiate(env, case):
    """Declare the case's parameters on `env` and build its scope.  Returns (scope, ins, outs)."""
    ins = [env.static_parameter(shape, name) for shape, name, _ in case.inputs]
    outs = [env.static_parameter(shape, name) for shape, name, _ in case.outputs]
    scope = env.scope()
    case.build(scope, ins, outs)
    return scope, ins, outs
