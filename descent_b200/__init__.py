"""descent_b200 -- Python mirror of descent's host API over the C ABI of libdescent_cuda.so.

Names and argument meaning follow the reference (sjb3d/descent): `Environment` (src/environment.rs),
`Scope` / `Array` / `UArray` / `DualArray` (src/array.rs), modules (src/module.rs), loss (src/loss.rs),
optimisers (src/optimizer.rs), so tests read like the reference's own (src/lib.rs:26-231).
Everything numeric happens in the shared library's CUDA kernels; this file only marshals handles.
Importing it without the built library raises: there is no Python or CPU execution path.
"""
import ctypes
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdescent_cuda.so")

INIT_ZERO, INIT_RAND_NORMAL, INIT_RAND_UNIFORM = 0, 1, 2


class DescentError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise DescentError(
            "%s is missing: run `python -m descent_b200.build` (or __graft_entry__.build()) first; "
            "descent_b200 has no fallback path" % LIB_PATH)
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()
lib.dsc_last_error.restype = ctypes.c_char_p
c_int_p = ctypes.POINTER(ctypes.c_int)
c_i64_p = ctypes.POINTER(ctypes.c_int64)
c_f32_p = ctypes.POINTER(ctypes.c_float)


def _check(rc):
    if rc != 0:
        raise DescentError(lib.dsc_last_error().decode("utf-8", "replace"))


def _i64(values):
    values = [int(v) for v in values]
    return (ctypes.c_int64 * max(1, len(values)))(*values), len(values)


def _ints(values):
    values = [int(v) for v in values]
    return (ctypes.c_int * max(1, len(values)))(*values), len(values)


def _take_string(ptr):
    s = ctypes.cast(ptr, ctypes.c_char_p).value.decode("utf-8")
    lib.dsc_string_free(ptr)
    return s


def device_count():
    n = ctypes.c_int(0)
    lib.dsc_device_count(ctypes.byref(n))
    return n.value


def nvrtc_compile(source, options=("-fmad=false",)):
    """Compile CUDA C for sm_100a; works without a GPU.  Returns the cubin size in bytes."""
    opts = (ctypes.c_char_p * max(1, len(options)))(*[o.encode() for o in options])
    cubin = ctypes.c_void_p()
    size = ctypes.c_size_t(0)
    _check(lib.dsc_nvrtc_compile(source.encode(), opts, len(options), ctypes.byref(cubin), ctypes.byref(size)))
    lib.dsc_host_buffer_free(cubin)
    return size.value


class Parameter:
    def __init__(self, env, pid):
        self.env, self.id = env, pid

    def _info(self):
        shape = (ctypes.c_int64 * 7)()
        ndim = ctypes.c_int(0)
        name = ctypes.create_string_buffer(64)
        trainable = ctypes.c_int(0)
        _check(lib.dsc_env_parameter_info(self.env._h, self.id, shape, ctypes.byref(ndim), name, ctypes.byref(trainable)))
        return tuple(shape[i] for i in range(ndim.value)), name.value.decode(), bool(trainable.value)

    def shape(self):
        return self._info()[0]

    def name(self):
        return self._info()[1]

    def is_trainable(self):
        return self._info()[2]

    def element_count(self):
        return int(np.prod(self.shape()))


class Graph:
    def __init__(self, handle, owned=True):
        self._h, self._owned = handle, owned

    def export_json(self):
        out = ctypes.c_void_p()
        _check(lib.dsc_graphdef_export_json(self._h, ctypes.byref(out)))
        return json.loads(_take_string(out))

    def kernel_source(self, sm_count=148, dp_rank=0, tf32=False):
        out = ctypes.c_void_p()
        _check(lib.dsc_graphdef_kernel_source_ex(self._h, sm_count, dp_rank, int(tf32), ctypes.byref(out)))
        return _take_string(out)

    def write_dot_file(self, mode, path):
        _check(lib.dsc_graphdef_write_dot_file(self._h, {"none": 0, "cluster": 1, "color": 2}[mode], path.encode()))

    def __del__(self):
        if getattr(self, "_owned", False) and self._h and lib is not None:  # `lib` is None while the interpreter shuts down
            lib.dsc_graphdef_destroy(self._h)
            self._h = None


class Environment:
    """Environment::new (environment.rs:104).  device=-1 gives a host-only environment for graph building."""

    def __init__(self, device=0):
        h = ctypes.c_void_p()
        _check(lib.dsc_env_create(device, ctypes.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib.dsc_env_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def ctx(self):
        c = ctypes.c_void_p()
        _check(lib.dsc_env_ctx(self._h, ctypes.byref(c)))
        return c

    def static_parameter(self, shape, name):
        s, n = _i64(shape)
        pid = ctypes.c_int(-1)
        _check(lib.dsc_env_static_parameter(self._h, s, n, name.encode(), ctypes.byref(pid)))
        return Parameter(self, pid.value)

    def trainable_parameter(self, shape, name, init_kind=INIT_ZERO, init_scale=0.0):
        s, n = _i64(shape)
        pid = ctypes.c_int(-1)
        _check(lib.dsc_env_trainable_parameter(self._h, s, n, name.encode(), init_kind, ctypes.c_float(init_scale), ctypes.byref(pid)))
        return Parameter(self, pid.value)

    def static_parameter_with_data(self, shape, name, data):
        p = self.static_parameter(shape, name)
        self.write(p, data)
        return p

    def parameter(self, pid):
        return Parameter(self, pid)

    def parameter_count(self):
        n = ctypes.c_int(0)
        _check(lib.dsc_env_parameter_count(self._h, ctypes.byref(n)))
        return n.value

    def write(self, param, data, pinned=False):
        """ParameterWriter: writes the floats from the start and zero-fills the rest (staging.rs:181-187)."""
        a = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
        _check(lib.dsc_env_write_parameter(self._h, param.id, a.ctypes.data_as(c_f32_p), ctypes.c_size_t(a.size), 1 if pinned else 0))

    def prefetch_pinned(self, param, pinned_array):
        """Start copying the next batch while earlier runs execute; it lands in `param` at the next run()."""
        _check(lib.dsc_env_prefetch_parameter(self._h, param.id, pinned_array.ctypes.data_as(c_f32_p), ctypes.c_size_t(pinned_array.size)))

    def write_pinned(self, param, pinned_array):
        _check(lib.dsc_env_write_parameter(self._h, param.id, pinned_array.ctypes.data_as(c_f32_p), ctypes.c_size_t(pinned_array.size), 1))

    def zero_fill(self, param):
        _check(lib.dsc_env_write_parameter(self._h, param.id, None, ctypes.c_size_t(0), 0))

    def read_parameter_to_vec(self, param):
        out = np.empty(param.element_count(), dtype=np.float32)
        _check(lib.dsc_env_read_parameter(self._h, param.id, out.ctypes.data_as(c_f32_p), ctypes.c_size_t(out.size)))
        return out

    def read(self, param):
        return self.read_parameter_to_vec(param).reshape(param.shape())

    def read_parameter_scalar(self, param):
        out = np.empty(1, dtype=np.float32)
        _check(lib.dsc_env_read_parameter(self._h, param.id, out.ctypes.data_as(c_f32_p), ctypes.c_size_t(1)))
        return float(out[0])

    def reset_parameter(self, param, rng_state):
        st = ctypes.c_uint64(rng_state)
        _check(lib.dsc_env_reset_parameter(self._h, param.id, ctypes.byref(st)))
        return st.value

    def scope(self):
        h = ctypes.c_void_p()
        _check(lib.dsc_env_scope(self._h, ctypes.byref(h)))
        return Scope(self, h)

    def build_graph(self, f):
        scope = self.scope()
        f(scope)
        return scope.build_graph()

    def run(self, graph, rand_seed):
        _check(lib.dsc_env_run(self._h, graph._h, ctypes.c_uint32(rand_seed & 0xFFFFFFFF)))

    def sync(self):
        _check(lib.dsc_env_sync(self._h))

    def set_options(self, use_cuda_graph=True, profile_runs=False):
        _check(lib.dsc_env_set_options(self._h, int(use_cuda_graph), int(profile_runs)))

    def set_tf32(self, on):
        """Tensor-core (tcgen05, TF32 operands, FP32 accumulate) path for plain dense MatMuls; default off = strict FP32."""
        _check(lib.dsc_env_set_tf32(self._h, int(on)))

    def set_sm_count(self, sm_count):
        """Plan as if the device had `sm_count` SMs (0 = real): persistent kernels walk many tiles per CTA (tests)."""
        _check(lib.dsc_env_set_sm_count(self._h, int(sm_count)))

    def print_timings(self, label):
        _check(lib.dsc_env_print_timings(self._h, label.encode()))

    def init_data_parallel(self, world, rank, unique_id):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        _check(lib.dsc_env_init_data_parallel(self._h, world, rank, buf))

    def set_data_parallel_for_tracing(self, world, rank):
        _check(lib.dsc_env_set_data_parallel_for_tracing(self._h, world, rank))

    def profile(self, graph, rand_seed=0, iterations=5):
        out = ctypes.c_void_p()
        _check(lib.dsc_env_profile(self._h, graph._h, ctypes.c_uint32(rand_seed), iterations, ctypes.byref(out)))
        return json.loads(_take_string(out))

    def graph_stats(self, graph):
        out = ctypes.c_void_p()
        _check(lib.dsc_env_graph_stats(self._h, graph._h, ctypes.byref(out)))
        return json.loads(_take_string(out))

    def example(self, network, mini_batch_size, optimizer="adam", weight_decay=1.0e-8, image_width=0, image_height=0):
        return Example(self, network, mini_batch_size, optimizer, weight_decay, image_width, image_height)


def nccl_unique_id():
    buf = ctypes.create_string_buffer(128)
    _check(lib.dsc_dp_unique_id(buf))
    return buf.raw


def pinned_array(count):
    """A float32 numpy array backed by pinned host memory (dsc_host_alloc)."""
    p = ctypes.c_void_p()
    _check(lib.dsc_host_alloc(ctypes.c_size_t(count * 4), ctypes.byref(p)))
    arr = np.ctypeslib.as_array(ctypes.cast(p, c_f32_p), shape=(count,))
    return arr


class Scope:
    def __init__(self, env, handle):
        self.env, self._h = env, handle

    def __del__(self):
        if self._h and lib is not None:
            lib.dsc_scope_destroy(self._h)
            self._h = None

    def _op(self, op, nodes, iargs=(), fargs=()):
        n, nn = _ints(nodes)
        ia, ni = _i64(iargs)
        fa = (ctypes.c_float * max(1, len(fargs)))(*fargs)
        out = (ctypes.c_int * 2)()
        nout = ctypes.c_int(0)
        _check(lib.dsc_array_op(self._h, op.encode(), n, nn, ia, ni, fa, len(fargs), out, ctypes.byref(nout)))
        return [out[i] for i in range(nout.value)]

    def _pair(self, fn, *args):
        v, g = ctypes.c_int(-1), ctypes.c_int(-1)
        _check(fn(self._h, *args, ctypes.byref(v), ctypes.byref(g)))
        return DualArray(Array(self, v.value), Array(self, g.value))

    def literal(self, value):
        return self._pair(lib.dsc_scope_literal, ctypes.c_float(value))

    def literal_u32(self, value):
        n = ctypes.c_int(-1)
        _check(lib.dsc_scope_literal_u32(self._h, ctypes.c_uint32(value), ctypes.byref(n)))
        return UArray(self, n.value)

    def coord(self, length):
        return self._pair(lib.dsc_scope_coord, ctypes.c_int64(length))

    def rand(self, shape):
        s, n = _i64(shape)
        return self._pair(lib.dsc_scope_rand, s, n)

    def parameter(self, param):
        return self._pair(lib.dsc_scope_parameter, param.id)

    def parameter_value(self, param):
        n = ctypes.c_int(-1)
        _check(lib.dsc_scope_parameter_value(self._h, param.id, ctypes.byref(n)))
        return Array(self, n.value)

    def write_parameter_value(self, param, rhs):
        _check(lib.dsc_scope_write_parameter_value(self._h, param.id, rhs.node))

    def update_parameter_value(self, param, f):
        result = f(self.parameter_value(param))
        self.write_parameter_value(param, result)
        return result

    def accumulator(self, shape):
        s, n = _i64(shape)
        node = ctypes.c_int(-1)
        _check(lib.dsc_scope_accumulator(self._h, s, n, ctypes.byref(node)))
        return Array(self, node.value)

    def next_colour(self):
        _check(lib.dsc_scope_next_colour(self._h))

    def trainable_parameters(self):
        buf = (ctypes.c_int * 256)()
        n = ctypes.c_int(0)
        _check(lib.dsc_scope_trainable_parameters(self._h, buf, 256, ctypes.byref(n)))
        return [Parameter(self.env, buf[i]) for i in range(n.value)]

    def all_reduce_gradients(self, params):
        p, n = _ints([q.id for q in params])
        _check(lib.dsc_scope_all_reduce_gradients(self._h, p, n))

    def build_graph(self):
        h = ctypes.c_void_p()
        _check(lib.dsc_scope_build_graph(self._h, ctypes.byref(h)))
        return Graph(h)

    def export_json(self):
        out = ctypes.c_void_p()
        _check(lib.dsc_scope_export_json(self._h, ctypes.byref(out)))
        return json.loads(_take_string(out))

    def _into_array(self, x):
        if isinstance(x, Array):
            return x
        if isinstance(x, Parameter):
            return self.parameter_value(x)
        if isinstance(x, DualArray):
            raise TypeError("expected an Array, got a DualArray (use .value())")
        return self.literal(float(x)).value()

    def _into_uarray(self, x):
        if isinstance(x, UArray):
            return x
        return self.literal_u32(int(x))

    def _into_dual(self, x):
        if isinstance(x, DualArray):
            return x
        if isinstance(x, Parameter):
            return self.parameter(x)
        return self.literal(float(x))


class _ArrayCommon:
    def __init__(self, scope, node):
        self.scope, self.node = scope, node

    def shape(self):
        shape = (ctypes.c_int64 * 7)()
        ndim = ctypes.c_int(0)
        _check(lib.dsc_array_shape(self.scope._h, self.node, shape, ctypes.byref(ndim)))
        return tuple(shape[i] for i in range(ndim.value))

    def _new(self, node):
        return type(self)(self.scope, node)

    def broadcast(self, shape):
        return self._new(self.scope._op("broadcast", [self.node], shape)[0])

    def limit_axis(self, axis, start, end):
        return self._new(self.scope._op("limit_axis", [self.node], [axis, start, end])[0])

    def lock_axis(self, axis, coord, keep_axis):
        return self._new(self.scope._op("lock_axis", [self.node], [axis, coord, int(keep_axis)])[0])

    def reshape(self, shape):
        return self._new(self.scope._op("reshape", [self.node], shape)[0])

    def transpose(self):
        return self._new(self.scope._op("transpose", [self.node])[0])


class UArray(_ArrayCommon):
    def _bin(self, op, rhs):
        return UArray(self.scope, self.scope._op(op, [self.node, self.scope._into_uarray(rhs).node])[0])

    def __add__(self, rhs):
        return self._bin("uadd", rhs)

    def __mul__(self, rhs):
        return self._bin("umul", rhs)

    def __mod__(self, rhs):
        return self._bin("urem", rhs)

    def __xor__(self, rhs):
        return self._bin("uxor", rhs)

    def to_f32_bits(self):
        return Array(self.scope, self.node)

    def into_f32(self):
        return Array(self.scope, self.scope._op("into_f32", [self.node])[0])


class Array(_ArrayCommon):
    def _bin(self, op, rhs, swap=False):
        r = self.scope._into_array(rhs)
        nodes = [r.node, self.node] if swap else [self.node, r.node]
        return Array(self.scope, self.scope._op(op, nodes)[0])

    def _un(self, op, iargs=()):
        return Array(self.scope, self.scope._op(op, [self.node], iargs)[0])

    def __add__(self, rhs):
        return self._bin("add", rhs)

    def __radd__(self, lhs):
        return self._bin("add", lhs, swap=True)

    def __sub__(self, rhs):
        return self._bin("sub", rhs)

    def __rsub__(self, lhs):
        return self._bin("sub", lhs, swap=True)

    def __mul__(self, rhs):
        return self._bin("mul", rhs)

    def __rmul__(self, lhs):
        return self._bin("mul", lhs, swap=True)

    def __truediv__(self, rhs):
        return self._bin("div", rhs)

    def __rtruediv__(self, lhs):
        return self._bin("div", lhs, swap=True)

    def __neg__(self):
        return self._un("neg")

    def concat(self, other, axis):
        return Array(self.scope, self.scope._op("concat", [self.node, self.scope._into_array(other).node], [axis])[0])

    def one_hot(self, count):
        return self._un("one_hot", [count])

    def reduce_max(self, axis, keep_axis):
        return self._un("reduce_max", [axis, int(keep_axis)])

    def reduce_sum(self, axis, keep_axis):
        return self._un("reduce_sum", [axis, int(keep_axis)])

    def argmax(self, axis, keep_axis):
        return self._un("argmax", [axis, int(keep_axis)])

    def coord(self, axis):
        return self._un("coord", [axis])

    def gather(self, axis, indices):
        return Array(self.scope, self.scope._op("gather", [self.node, self.scope._into_uarray(indices).node], [axis])[0])

    def scatter_add(self, values, axis, indices):
        nodes = [self.node, self.scope._into_array(values).node, self.scope._into_uarray(indices).node]
        return Array(self.scope, self.scope._op("scatter_add", nodes, [axis])[0])

    def _select(self, op, rhs, passed, failed):
        s = self.scope
        nodes = [self.node, s._into_array(rhs).node, s._into_array(passed).node, s._into_array(failed).node]
        return Array(s, s._op(op, nodes)[0])

    def select_eq(self, rhs, passed, failed):
        return self._select("select_eq", rhs, passed, failed)

    def select_gt(self, rhs, passed, failed):
        return self._select("select_gt", rhs, passed, failed)

    def square(self):
        return self._un("square")

    def sqrt(self):
        return self._un("sqrt")

    def exp(self):
        return self._un("exp")

    def log(self):
        return self._un("log")

    def sin(self):
        return self._un("sin")

    def cos(self):
        return self._un("cos")

    def sigmoid(self):
        return self._un("sigmoid")

    def tanh(self):
        return self._un("tanh")

    def to_u32_bits(self):
        return UArray(self.scope, self.node)

    def into_u32(self):
        return UArray(self.scope, self.scope._op("into_u32", [self.node])[0])

    def pow(self, rhs):
        return self._bin("pow", rhs)

    def matmul(self, rhs):
        return self._bin("matmul", rhs)

    def accumulate(self, src):
        self.scope._op("accumulate", [self.node, self.scope._into_array(src).node])

    def pad_image(self, pad):
        return self._un("pad_image", [pad])

    def unpad_image(self, pad):
        return self._un("unpad_image", [pad])

    def with_empty_grad(self):
        return self, self.scope.accumulator(self.shape())


class DualArray:
    def __init__(self, value, loss_grad):
        self._value, self._grad, self.scope = value, loss_grad, value.scope

    def value(self):
        return self._value

    def loss_grad(self):
        return self._grad

    def into_inner(self):
        return self._value, self._grad

    def shape(self):
        return self._value.shape()

    def _nodes(self):
        return [self._value.node, self._grad.node]

    def _ret(self, out):
        return DualArray(Array(self.scope, out[0]), Array(self.scope, out[1]))

    def _un(self, op, iargs=(), fargs=()):
        return self._ret(self.scope._op(op, self._nodes(), iargs, fargs))

    def _bin(self, op, rhs, iargs=()):
        return self._ret(self.scope._op(op, self._nodes() + self.scope._into_dual(rhs)._nodes(), iargs))

    def __add__(self, rhs):
        return self._bin("dual.add", rhs)

    def __sub__(self, rhs):
        return self._bin("dual.sub", rhs)

    def __mul__(self, rhs):
        return self._bin("dual.mul", rhs)

    def square(self):
        return self._un("dual.square")

    def sin(self):
        return self._un("dual.sin")

    def tanh(self):
        return self._un("dual.tanh")

    def sigmoid(self):
        return self._un("dual.sigmoid")

    def leaky_relu(self, leakiness):
        return self._un("dual.leaky_relu", (), [leakiness])

    def matmul(self, rhs):
        return self._bin("dual.matmul", rhs)

    def transpose(self):
        return self._un("dual.transpose")

    def pow(self, rhs):
        return self._bin("dual.pow", rhs)

    def select_eq(self, rhs, passed, failed):
        s = self.scope
        nodes = self._nodes() + s._into_dual(rhs)._nodes() + s._into_dual(passed)._nodes() + s._into_dual(failed)._nodes()
        return self._ret(s._op("dual.select_eq", nodes))

    def lock_axis(self, axis, coord, keep_axis):
        return self._un("dual.lock_axis", [axis, coord, int(keep_axis)])

    def reshape(self, shape):
        return self._un("dual.reshape", shape)

    def next_colour(self):
        self.scope.next_colour()
        return self

    def map(self, f):
        return f(self)

    def apply(self, module, is_training):
        return module.eval(self, is_training)

    def conv2d(self, filt, pad, stride):
        return self._bin("dual.conv2d", filt, [pad, stride[0], stride[1]])

    def max_pool2d(self, filt, stride):
        return self._un("dual.max_pool2d", [filt[0], filt[1], stride[0], stride[1]])

    def reduce_sum(self, axis, keep_axis):
        return self._un("dual.reduce_sum", [axis, int(keep_axis)])

    def reduce_max(self, axis, keep_axis):
        return self._un("dual.reduce_max", [axis, int(keep_axis)])

    def flatten(self):
        return self._un("dual.flatten")

    def set_loss(self):
        return Array(self.scope, self.scope._op("dual.set_loss", self._nodes())[0])

    def concat(self, other, axis):
        return self._bin("dual.concat", other, [axis])


class Module:
    def __init__(self, env, handle):
        self.env, self._h = env, handle

    def eval(self, x, is_training):
        v, g = ctypes.c_int(-1), ctypes.c_int(-1)
        _check(lib.dsc_module_eval(self.env._h, x.scope._h, self._h, x.value().node, x.loss_grad().node, int(is_training),
                                   ctypes.byref(v), ctypes.byref(g)))
        return DualArray(Array(x.scope, v.value), Array(x.scope, g.value))

    def train(self, x):
        return self.eval(x, True)

    def test(self, x):
        return self.eval(x, False)


def _module(env, fn, *args):
    h = ctypes.c_int(-1)
    _check(fn(env._h, *args, ctypes.byref(h)))
    return Module(env, h.value)


def Dense(env, input, output, w_init=None, b_init=None):
    wk, ws = w_init if w_init else (-1, 0.0)
    bk, bs = b_init if b_init else (-1, 0.0)
    return _module(env, lib.dsc_module_dense, ctypes.c_int64(input), ctypes.c_int64(output), wk, ctypes.c_float(ws), bk, ctypes.c_float(bs))


def Conv2D(env, input_channels, output_channels, filter_w, filter_h, pad=0, stride=(1, 1), groups=1, blur=False):
    i64 = ctypes.c_int64
    return _module(env, lib.dsc_module_conv2d, i64(input_channels), i64(output_channels), i64(filter_w), i64(filter_h), i64(pad),
                   i64(stride[0]), i64(stride[1]), i64(groups), int(blur))


def MaxPool2D(env):
    return _module(env, lib.dsc_module_max_pool2d)


def MaxBlurPool2D(env, channels):
    return _module(env, lib.dsc_module_max_blur_pool2d, ctypes.c_int64(channels))


def Dropout(env, amount):
    return _module(env, lib.dsc_module_dropout, ctypes.c_float(amount))


def LSTMCell(env, input, output):
    return _module(env, lib.dsc_module_lstm_cell, ctypes.c_int64(input), ctypes.c_int64(output))


def softmax_cross_entropy_loss(z, y):
    s = z.scope
    v, g = ctypes.c_int(-1), ctypes.c_int(-1)
    _check(lib.dsc_softmax_cross_entropy_loss(s._h, z.value().node, z.loss_grad().node, s._into_array(y).node, ctypes.byref(v), ctypes.byref(g)))
    return DualArray(Array(s, v.value), Array(s, g.value))


def softmax_cross_entropy_accuracy(z, y):
    s = z.scope
    n = ctypes.c_int(-1)
    _check(lib.dsc_softmax_cross_entropy_accuracy(s._h, z.value().node, z.loss_grad().node, s._into_array(y).node, ctypes.byref(n)))
    return Array(s, n.value)


def add_weight_decay_to_grad(scope, parameters, weight_decay):
    p, n = _ints([q.id for q in parameters])
    _check(lib.dsc_add_weight_decay_to_grad(scope._h, p, n, ctypes.c_float(weight_decay)))


class Optimizer:
    def __init__(self, env, handle):
        self.env, self._h = env, handle

    def reset_state(self):
        _check(lib.dsc_optimizer_reset_state(self.env._h, self._h))

    def state(self):
        buf = (ctypes.c_int * 512)()
        n = ctypes.c_int(0)
        _check(lib.dsc_optimizer_state(self.env._h, self._h, buf, 512, ctypes.byref(n)))
        return [Parameter(self.env, buf[i]) for i in range(n.value)]


def StochasticGradientDescent(env, scope, parameters, learning_rate, momentum):
    p, n = _ints([q.id for q in parameters])
    h = ctypes.c_int(-1)
    _check(lib.dsc_optimizer_sgd(env._h, scope._h, p, n, scope._into_array(learning_rate).node, ctypes.c_float(momentum), ctypes.byref(h)))
    return Optimizer(env, h.value)


def Adam(env, scope, parameters, learning_rate, beta1, beta2, epsilon):
    p, n = _ints([q.id for q in parameters])
    h = ctypes.c_int(-1)
    _check(lib.dsc_optimizer_adam(env._h, scope._h, p, n, scope._into_array(learning_rate).node, ctypes.c_float(beta1), ctypes.c_float(beta2),
                                  ctypes.c_float(epsilon), ctypes.byref(h)))
    return Optimizer(env, h.value)


class _ExampleStruct(ctypes.Structure):
    _fields_ = [("x", ctypes.c_int), ("y", ctypes.c_int), ("learning_rate_scale", ctypes.c_int), ("loss_sum", ctypes.c_int),
                ("accuracy_sum", ctypes.c_int), ("image", ctypes.c_int), ("num_parameters", ctypes.c_int), ("parameters", ctypes.c_int * 64),
                ("num_optimizer_state", ctypes.c_int), ("optimizer_state", ctypes.c_int * 130), ("train_graph", ctypes.c_void_p),
                ("test_graph", ctypes.c_void_p)]


class Example:
    """One of the reference's example networks with its training / test graphs (descent_b200/csrc/examples.cpp)."""

    def __init__(self, env, network, mini_batch_size, optimizer, weight_decay, image_width, image_height):
        st = _ExampleStruct()
        _check(lib.dsc_example_create(env._h, network.encode(), ctypes.c_int64(mini_batch_size), optimizer.encode(), ctypes.c_float(weight_decay),
                                      ctypes.c_int64(image_width), ctypes.c_int64(image_height), ctypes.byref(st)))
        self.env, self.network, self.mini_batch_size = env, network, mini_batch_size

        def P(i):
            return Parameter(env, i) if i >= 0 else None
        self.x, self.y, self.learning_rate_scale = P(st.x), P(st.y), P(st.learning_rate_scale)
        self.loss_sum, self.accuracy_sum, self.image = P(st.loss_sum), P(st.accuracy_sum), P(st.image)
        self.parameters = [Parameter(env, st.parameters[i]) for i in range(st.num_parameters)]
        self.optimizer_state = [Parameter(env, st.optimizer_state[i]) for i in range(st.num_optimizer_state)]
        self.train_graph = Graph(ctypes.c_void_p(st.train_graph), owned=True)
        self.test_graph = Graph(ctypes.c_void_p(st.test_graph), owned=True) if st.test_graph else None
        self.train_graph_json = self._json(0)
        self.test_graph_json = self._json(1) if st.test_graph else None

    def _json(self, which):
        out = ctypes.c_void_p()
        _check(lib.dsc_example_graph_json(self.env._h, which, ctypes.byref(out)))
        return json.loads(_take_string(out))


# ---- host random numbers and data front ends of the reference's examples ---------------
c_u8_p = ctypes.POINTER(ctypes.c_uint8)
c_u64_p = ctypes.POINTER(ctypes.c_uint64)


class ChaCha20Rng:
    """`rand_chacha::ChaCha20Rng::seed_from_u64(seed)` with the rand 0.8 sampling rules the examples use
    (descent_b200/csrc/host_rng.hpp): next_u32 (the per-step rand_seed), Open01 floats, gen_range, shuffle."""

    def __init__(self, seed):
        self._h = ctypes.c_void_p()
        _check(lib.dsc_rng_create(ctypes.c_uint64(seed), ctypes.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.dsc_rng_destroy(self._h)
            self._h = None

    def next_u32(self):
        out = ctypes.c_uint32()
        _check(lib.dsc_rng_next_u32(self._h, ctypes.byref(out)))
        return out.value

    def next_u64(self):
        out = ctypes.c_uint64()
        _check(lib.dsc_rng_next_u64(self._h, ctypes.byref(out)))
        return out.value

    def open01(self, count):
        out = np.empty(count, dtype=np.float32)
        _check(lib.dsc_rng_open01_f32(self._h, out.ctypes.data_as(c_f32_p), ctypes.c_size_t(count)))
        return out

    def gen_range(self, low, high, u32=False):
        """rng.gen_range(low..high): `u32=True` for u32 bounds, otherwise usize (64-bit draws) as in image_fit/main.rs:377."""
        out = ctypes.c_uint64()
        _check(lib.dsc_rng_gen_range(self._h, ctypes.c_uint64(low), ctypes.c_uint64(high), int(u32), ctypes.byref(out)))
        return out.value

    def gen_range_pairs(self, high0, high1, pairs):
        """`pairs` draws of (gen_range(0..high0), gen_range(0..high1)) as usize, in that order (image_fit/main.rs:376-378): [pairs, 2] uint64."""
        out = np.empty((pairs, 2), dtype=np.uint64)
        _check(lib.dsc_rng_gen_range_pairs(self._h, ctypes.c_uint64(high0), ctypes.c_uint64(high1), out.ctypes.data_as(c_u64_p), ctypes.c_size_t(pairs)))
        return out

    def shuffle(self, indices):
        """indices.shuffle(&mut rng) (fashion_mnist/main.rs:382): returns the shuffled copy as uint64."""
        arr = np.ascontiguousarray(indices, dtype=np.uint64).copy()
        _check(lib.dsc_rng_shuffle(self._h, arr.ctypes.data_as(c_u64_p), ctypes.c_size_t(arr.size)))
        return arr


def _reset_parameter_rng(self, param, rng):
    """Environment::reset_parameter(param, &mut rng) with the reference's generator (environment.rs:190-202)."""
    _check(lib.dsc_env_reset_parameter_rng(self._h, param.id, rng._h))


Environment.reset_parameter_rng = _reset_parameter_rng


def _take_bytes(ptr, size):
    data = ctypes.string_at(ptr, size)
    lib.dsc_free_bytes(ptr)
    return data


def load_gz_bytes(path):
    """examples/fashion_mnist/main.rs:13-19."""
    ptr, size = c_u8_p(), ctypes.c_size_t()
    _check(lib.dsc_load_gz_bytes(os.fsencode(path), ctypes.byref(ptr), ctypes.byref(size)))
    return _take_bytes(ptr, size.value)


def gunzip(data):
    ptr, size = c_u8_p(), ctypes.c_size_t()
    _check(lib.dsc_gunzip(data, ctypes.c_size_t(len(data)), ctypes.byref(ptr), ctypes.byref(size)))
    return _take_bytes(ptr, size.value)


def read_images_info(data):
    """(images, rows, cols) of an IDX image file (fashion_mnist/main.rs:26-33)."""
    n, r, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    _check(lib.dsc_idx_images_info(data, ctypes.c_size_t(len(data)), ctypes.byref(n), ctypes.byref(r), ctypes.byref(c)))
    return n.value, r.value, c.value


def read_labels_info(data):
    n = ctypes.c_uint32()
    _check(lib.dsc_idx_labels_info(data, ctypes.c_size_t(len(data)), ctypes.byref(n)))
    return n.value


def unpack_images(data, indices):
    """[len(indices), rows * cols] float32 = byte / 255 (fashion_mnist/main.rs:42-60)."""
    _, rows, cols = read_images_info(data)
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    out = np.empty((idx.size, rows * cols), dtype=np.float32)
    _check(lib.dsc_idx_unpack_images(data, ctypes.c_size_t(len(data)), idx.ctypes.data_as(c_u64_p), ctypes.c_size_t(idx.size), out.ctypes.data_as(c_f32_p)))
    return out


def unpack_labels(data, indices):
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    out = np.empty(idx.size, dtype=np.float32)
    _check(lib.dsc_idx_unpack_labels(data, ctypes.c_size_t(len(data)), idx.ctypes.data_as(c_u64_p), ctypes.c_size_t(idx.size), out.ctypes.data_as(c_f32_p)))
    return out


def decode_jpeg_rgb(data):
    """[height, width, 3] uint8 of a baseline JPEG file (image_fit/main.rs:278-282)."""
    w, h, ptr = ctypes.c_int(), ctypes.c_int(), c_u8_p()
    _check(lib.dsc_jpeg_decode_rgb(data, ctypes.c_size_t(len(data)), ctypes.byref(w), ctypes.byref(h), ctypes.byref(ptr)))
    raw = _take_bytes(ptr, w.value * h.value * 3)
    return np.frombuffer(raw, dtype=np.uint8).reshape(h.value, w.value, 3).copy()


def write_ppm(path, rgb):
    """A predicted image [height, width, 3] in [0, 1] as a binary PPM (byte = x * 255 + 0.5 clamped, image_fit/main.rs:423-426)."""
    arr = np.ascontiguousarray(rgb, dtype=np.float32)
    _check(lib.dsc_write_ppm(os.fsencode(path), arr.ctypes.data_as(c_f32_p), arr.shape[1], arr.shape[0]))


def write_csv_row(stream, values):
    """One row of the examples' statistics files (fashion_mnist/main.rs:398-412, image_fit/main.rs:408-416): `"label", v, v, ...`."""
    stream.write(", ".join('"%s"' % v if isinstance(v, str) else repr(float(v)) if isinstance(v, float) else str(v) for v in values) + "\n")


# ---- op-level entry points (include/descent_api.h "Op-level entry points", csrc/ops_api.cpp) -------------------------------
class Op:
    """One fused op planned for a shape (dsc_op_*): owns its operand / result buffers on the device."""

    def __init__(self, env, handle):
        self.env, self._h = env, handle

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.dsc_op_destroy(self._h)
            self._h = None

    def parameter(self, name):
        pid = ctypes.c_int()
        _check(lib.dsc_op_parameter(self._h, name.encode(), ctypes.byref(pid)))
        return self.env.parameter(pid.value)

    def device_buffer(self, name):
        """(device address, bytes) of a buffer: what a caller's own kernels would read / write in place."""
        ptr, size = ctypes.c_void_p(), ctypes.c_size_t()
        _check(lib.dsc_op_buffer(self._h, name.encode(), ctypes.byref(ptr), ctypes.byref(size)))
        return ptr.value, size.value

    def write(self, name, data):
        self.env.write(self.parameter(name), data)

    def read(self, name):
        return self.env.read(self.parameter(name))

    def run(self, rand_seed=0):
        _check(lib.dsc_op_run(self._h, ctypes.c_uint32(rand_seed)))


def _op_conv2d(self, images, height, width, in_channels, out_channels, filter_h, filter_w, pad=0, stride=(1, 1), groups=1, backward=False):
    h = ctypes.c_void_p()
    _check(lib.dsc_op_conv2d(self._h, *(ctypes.c_int64(v) for v in (images, height, width, in_channels, out_channels, filter_h, filter_w, pad, stride[0], stride[1], groups)),
                             int(backward), ctypes.byref(h)))
    return Op(self, h)


def _op_scatter_add(self, rows, inner, count):
    h = ctypes.c_void_p()
    _check(lib.dsc_op_scatter_add(self._h, ctypes.c_int64(rows), ctypes.c_int64(inner), ctypes.c_int64(count), ctypes.byref(h)))
    return Op(self, h)


def _op_softmax_cross_entropy(self, rows, classes):
    h = ctypes.c_void_p()
    _check(lib.dsc_op_softmax_cross_entropy(self._h, ctypes.c_int64(rows), ctypes.c_int64(classes), ctypes.byref(h)))
    return Op(self, h)


def _op_adam_step(self, counts, learning_rate, beta1=0.9, beta2=0.999, epsilon=1.0e-8):
    arr, n = _i64(counts)
    h = ctypes.c_void_p()
    _check(lib.dsc_op_adam_step(self._h, arr, n, ctypes.c_float(learning_rate), ctypes.c_float(beta1), ctypes.c_float(beta2), ctypes.c_float(epsilon), ctypes.byref(h)))
    return Op(self, h)


Environment.op_conv2d = _op_conv2d
Environment.op_scatter_add = _op_scatter_add
Environment.op_softmax_cross_entropy = _op_softmax_cross_entropy
Environment.op_adam_step = _op_adam_step
