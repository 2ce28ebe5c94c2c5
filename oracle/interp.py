"""numpy interpreter of descent's op graph (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Input: the JSON the frontend exports for a *raw* (pre-optimisation) graph: nodes in topological
order, each with its op, shape and argument edges; every edge carries a chain of views
(reference: View, src/shape.rs:354-360).  Values are flat float32 arrays; u32 values live in the same
storage as raw bits (kernel_common.glsl:218-220).
"""
import numpy as np

F32 = np.float32
U32 = np.uint32


# ---- integer helpers: bit-exact (kernel_common.glsl:205-216) --------------

def pcg(v):
    v = np.asarray(v, dtype=np.uint64) & 0xFFFFFFFF
    state = (v * 747796405 + 2891336453) & 0xFFFFFFFF
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
    return ((word >> 22) ^ word).astype(np.uint32)


def rand_from_index(uid, index, seed):
    """float(hash) / float(0xffffffffu): the divisor rounds to 2**32, so this is RNE(hash) * 2**-32."""
    h = pcg((pcg(index).astype(np.uint64) + np.uint64(seed) + np.uint64(uid)) & 0xFFFFFFFF)
    return (h.astype(np.float32) * np.float32(2.0 ** -32)).astype(np.float32)


def float_to_uint(x):
    """GLSL uint(float): truncation; negative/NaN are undefined in the reference (kernel.rs:293).
    The CUDA backend uses cvt.rzi.u32.f32 (saturating, NaN -> 0); the oracle pins the same choice."""
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    x = np.where(np.isnan(x), 0.0, x)
    return np.clip(np.trunc(x), 0.0, 4294967295.0).astype(np.uint64).astype(np.uint32)


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def from_bits(u):
    return np.ascontiguousarray(u, dtype=np.uint32).view(np.float32)


# ---- views (shape.rs:354-360, kernel.rs:89-133) -----------------------------------

def _strides(shape):
    s = [1] * len(shape)
    for i in range(len(shape) - 2, -1, -1):
        s[i] = s[i + 1] * shape[i + 1]
    return s


def view_index(view, idx):
    """Map linear indices in view.output_shape to linear indices in view.input_shape:
    coord_in[a] = clamp(offset[a] + sum(step * coord_out[i]), 0, len_a - 1)."""
    out_shape, in_shape = view["output_shape"], view["input_shape"]
    ostr, istr = _strides(out_shape), _strides(in_shape)
    coords_in = [np.full(idx.shape, off, dtype=np.int64) for off in view["input_offsets"]]
    for i, m in enumerate(view["mapping"]):
        if m is None:
            continue
        axis, step = m
        coords_in[axis] = coords_in[axis] + step * ((idx // ostr[i]) % out_shape[i])
    lin = np.zeros(idx.shape, dtype=np.int64)
    for a, c in enumerate(coords_in):
        lin += np.clip(c, 0, in_shape[a] - 1) * istr[a]
    return lin


def chain_index(chain, idx=None):
    if idx is None:
        idx = np.arange(chain["output_count"], dtype=np.int64)
    for view in reversed(chain["views"]):
        idx = view_index(view, idx)
    return idx


def apply_chain(src_flat, chain):
    if not chain["views"]:
        assert src_flat.size == chain["output_count"], (src_flat.size, chain)
        return src_flat
    return src_flat[chain_index(chain)]


# ---- ops ----------------------------------------------------------------------------------------

def _unary(kind, a):
    with np.errstate(all="ignore"):
        if kind == "Mov":
            return a
        if kind == "Neg":
            return -a
        if kind == "Sqrt":
            return np.sqrt(a, dtype=F32)
        if kind == "Exp":
            return np.exp(a, dtype=F32)
        if kind == "Log":
            return np.log(a, dtype=F32)
        if kind == "Sin":
            return np.sin(a.astype(np.float64)).astype(F32)  # correctly rounded: large SIREN arguments
        if kind == "Cos":
            return np.cos(a.astype(np.float64)).astype(F32)
        if kind == "FloatToUint":
            return from_bits(float_to_uint(a))
        if kind == "UintToFloat":
            return bits(a).astype(F32)
    raise ValueError(kind)


def _binary(kind, a, b):
    with np.errstate(all="ignore"):
        if kind == "Add":
            return (a + b).astype(F32)
        if kind == "Sub":
            return (a - b).astype(F32)
        if kind == "Mul":
            return (a * b).astype(F32)
        if kind == "Div":
            return (a / b).astype(F32)
        if kind == "Pow":
            return np.power(a.astype(np.float64), b.astype(np.float64)).astype(F32)
        ua, ub = bits(a).astype(np.uint64), bits(b).astype(np.uint64)
        if kind == "UAdd":
            return from_bits(((ua + ub) & 0xFFFFFFFF).astype(U32))
        if kind == "UMul":
            return from_bits(((ua * ub) & 0xFFFFFFFF).astype(U32))
        if kind == "URem":
            return from_bits((ua % np.maximum(ub, 1)).astype(U32))
        if kind == "UBitXor":
            return from_bits((ua ^ ub).astype(U32))
    raise ValueError(kind)


def _reduce(kind, a, arg_shape, axis):
    x = a.reshape(arg_shape)
    if kind == "Max":
        return np.max(x, axis=axis, keepdims=True).astype(F32).reshape(-1)  # kernel.rs:605,618
    return np.sum(x.astype(np.float64), axis=axis, keepdims=True).astype(F32).reshape(-1)  # kernel.rs:606,619


def tf32_operand(x, rounding):
    """FP32 -> TF32 operand as the tensor core sees it: 10 explicit mantissa bits.  `rounding`: "trunc" drops the
    low 13 bits, "rna" rounds to nearest (ties away).  Used only when a test asks for the TF32-emulating oracle."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    if rounding == "rna":
        u = u + np.uint32(0x1000)
    return (u & np.uint32(0xFFFFE000)).view(np.float32)


def _matmul(a, b, a_shape, b_shape, out_shape, mode, tf32=None):
    """array.rs:492-520, kernel.rs:385-557: [r, b, m, n] (Batches) or [r, m, b, n] (Rows); chunk c covers
    k in [c*chunk, (c+1)*chunk) with chunk = ceil(ceil(K/16)/r)*16 (kernel.rs:436)."""
    bc, m, k = a_shape
    _, _, n = b_shape
    r = out_shape[0]
    chunk = -(-(-(-k // 16)) // r) * 16
    if tf32:
        a, b = tf32_operand(a, tf32), tf32_operand(b, tf32)
    A = a.reshape(a_shape).astype(np.float64)
    B = b.reshape(b_shape).astype(np.float64)
    out = np.zeros((r, bc, m, n), dtype=F32)
    for c in range(r):
        lo, hi = c * chunk, min(k, (c + 1) * chunk)
        if lo < hi:
            out[c] = np.matmul(A[:, :, lo:hi], B[:, lo:hi, :]).astype(F32)
    if mode == "Rows":
        out = out.transpose(0, 2, 1, 3)
    return np.ascontiguousarray(out).reshape(-1)


def _unpad(a, arg_shape, axis, pad):
    """Adjoint of replicate padding along one axis (kernel.rs:670-690): edges sum their pad replicas."""
    x = np.moveaxis(a.reshape(arg_shape).astype(np.float64), axis, 0)
    n = x.shape[0] - 2 * pad
    out = x[pad:pad + n].copy()
    out[0] += x[:pad].sum(axis=0)
    out[n - 1] += x[pad + n:].sum(axis=0)
    return np.ascontiguousarray(np.moveaxis(out, 0, axis)).astype(F32).reshape(-1)


def _windows_to_image(a, arg_shape, out_shape, stride_w, stride_h):
    """col2im: the mathematically correct adjoint of image_to_windows (array.rs:559-600).  The
    reference kernel omits the window-range check (kernel.rs:770-787); decision recorded
    there: implement the correct adjoint."""
    lead = int(np.prod(arg_shape[:-6])) if len(arg_shape) > 6 else 1
    oh, ow, g, fh, fw, gnc = arg_shape[-6:]
    ih, iw, ic = out_shape[-3:]
    w = a.reshape(lead, oh, ow, g, fh, fw, gnc).astype(np.float64)
    img = np.zeros((lead, ih, iw, g, gnc), dtype=np.float64)
    for fy in range(fh):
        for fx in range(fw):
            img[:, fy:fy + (oh - 1) * stride_h + 1:stride_h, fx:fx + (ow - 1) * stride_w + 1:stride_w] += w[:, :, :, :, fy, fx, :]
    return img.reshape(lead, ih, iw, ic).astype(F32).reshape(-1)


def _gather(values, idx_bits, node_shape, values_shape, axis):
    """out[.., i, ..] = values[.., F2I(index), ..] (kernel.rs:336-351); index is the float's bit pattern."""
    v = values.reshape(values_shape)
    i = bits(idx_bits).astype(np.int64).reshape(node_shape)
    return np.take_along_axis(v, i, axis=axis).reshape(-1)


def _scatter_add(acc, values, idx_bits, shape, values_shape, axis):
    """acc[.., idx[i], ..] += values[.., i, ..] (kernel.rs:812-874); the reference's atomic order is
    unspecified, the oracle accumulates in float64 and rounds once."""
    out = np.moveaxis(acc.reshape(shape).astype(np.float64), axis, 0).copy()
    v = np.moveaxis(values.reshape(values_shape).astype(np.float64), axis, 0)
    i = bits(idx_bits).astype(np.int64)
    np.add.at(out, i, v)
    return np.ascontiguousarray(np.moveaxis(out, 0, axis)).astype(F32).reshape(-1)


class _Interp:
    def __init__(self, graph, params, rand_seed, dp_rank, tf32=None):
        self.graph, self.params, self.seed, self.rank, self.tf32 = graph, params, rand_seed, dp_rank, tf32
        self.nodes = graph["nodes"]
        self.by_id = {n["id"]: n for n in self.nodes}
        self.values = {}
        self.outputs = {}

    def tf32_mode_for(self, node_id):
        """`tf32` is None (strict), a rounding mode applied to every MatMul, or (mode, {node ids}) to emulate TF32 only
        on the MatMuls the backend actually ran on tensor cores."""
        if self.tf32 is None or isinstance(self.tf32, str):
            return self.tf32
        mode, nodes = self.tf32
        return mode if node_id in nodes else None

    def arg(self, node, k):
        e = node["args"][k]
        src = self.by_id[e["src"]]
        chain = e["chain"]
        if src["op"] == "Coord":  # float(linear index through the view), kernel.rs:266-270
            return chain_index(chain).astype(F32)
        if src["op"] == "Rand":  # kernel.rs:271-279; flat index of the unsharded tensor
            n = int(np.prod(src["shape"]))
            return rand_from_index(src["uid"], chain_index(chain) + self.rank * n, self.seed)
        return apply_chain(self.values[e["src"]], chain)

    def eval(self, node):
        op = node["op"]
        count = int(np.prod(node["shape"]))
        if op == "Input":
            v = np.ascontiguousarray(self.params[node["parameter"]], dtype=F32).reshape(-1)
            assert v.size == count, "parameter %d has %d elements, graph expects %d" % (node["parameter"], v.size, count)
            return v
        if op == "Literal":
            return from_bits(np.array([node["bits"]], dtype=U32))
        if op in ("Coord", "Rand"):
            return None  # evaluated at the consumer, through its view
        if op == "Output":
            self.outputs[node["parameter"]] = self.arg(node, 0).reshape(node["shape"]).copy()
            return None
        if op == "Unary":
            return _unary(node["kind"], self.arg(node, 0))
        if op == "Binary":
            return _binary(node["kind"], self.arg(node, 0), self.arg(node, 1))
        if op == "Select":
            a, b, p, f = (self.arg(node, k) for k in range(4))
            with np.errstate(invalid="ignore"):
                cond = (a == b) if node["kind"] == "Eq" else (a > b)
            return np.where(cond, p, f).astype(F32)
        if op == "Reduce":
            return _reduce(node["kind"], self.arg(node, 0), node["args"][0]["arg_shape"], node["axis"])
        if op == "MatMul":
            return _matmul(self.arg(node, 0), self.arg(node, 1), node["args"][0]["arg_shape"], node["args"][1]["arg_shape"], node["shape"], node["mode"],
                           self.tf32_mode_for(node["id"]))
        if op == "Unpad":
            return _unpad(self.arg(node, 0), node["args"][0]["arg_shape"], node["axis"], node["pad"])
        if op == "WindowsToImage":
            return _windows_to_image(self.arg(node, 0), node["args"][0]["arg_shape"], node["shape"], node["stride_w"], node["stride_h"])
        if op == "Gather":
            return _gather(self.arg(node, 0), self.arg(node, 1), node["shape"], node["args"][0]["arg_shape"], node["axis"])
        if op == "ScatterAdd":
            return _scatter_add(self.arg(node, 0), self.arg(node, 1), self.arg(node, 2), node["shape"], node["args"][1]["arg_shape"], node["axis"])
        raise ValueError("unknown op %r" % op)


def _live_nodes(graph):
    """Nodes an Output depends on (the reference's dead-code elimination, graph.rs:150-165): gradient
    accumulators that nothing feeds are legal as long as they are dead."""
    by_id = {n["id"]: n for n in graph["nodes"]}
    live, stack = set(), [n["id"] for n in graph["nodes"] if n["op"] == "Output"]
    while stack:
        i = stack.pop()
        if i in live:
            continue
        live.add(i)
        stack.extend(e["src"] for e in by_id[i]["args"])
    return live


def run_graph(graph, params, rand_seed=0, tf32=None):
    """Evaluate one run of `graph` (Environment::run, environment.rs:326-516).  `params`: {parameter id:
    array}.  Returns {parameter id: new value} for every Output.  `tf32` ("trunc" / "rna"): emulate TF32
    operand precision in every MatMul (for checking the tensor-core path); default None = strict FP32."""
    return run_graph_data_parallel([graph], [params], rand_seed, tf32)[0]


def run_graph_data_parallel(graphs, params_per_rank, rand_seed=0, tf32=None):
    """Lock-step evaluation of one graph per rank; AllReduce nodes sum their input over ranks
    (float64, rounded once).  With a single rank AllReduce is the identity."""
    interps = [_Interp(g, p, rand_seed, r, tf32) for r, (g, p) in enumerate(zip(graphs, params_per_rank))]
    live = _live_nodes(graphs[0])
    for pos, node0 in enumerate(graphs[0]["nodes"]):
        if node0["id"] not in live:
            continue
        if node0["op"] == "AllReduce":
            parts = [it.arg(it.nodes[pos], 0).astype(np.float64) for it in interps]
            total = np.sum(parts, axis=0).astype(F32)
            for it in interps:
                it.values[it.nodes[pos]["id"]] = total
            continue
        for it in interps:
            node = it.nodes[pos]
            assert node["op"] == node0["op"], "rank graphs differ in structure"
            if node["op"] == "Unary" and node["kind"] == "Mov" and not node["args"]:
                raise ValueError("live gradient accumulator %d was never written" % node["id"])
            v = it.eval(node)
            if v is not None:
                it.values[node["id"]] = v
    return [it.outputs for it in interps]
