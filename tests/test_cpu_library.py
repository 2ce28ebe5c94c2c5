"""CPU-side checks: the C-ABI library loads and exports every symbol the headers declare, generated kernels
compile for sm_100a with NVRTC (no GPU needed), the graph compiler reproduces the structure fixtures of the
reference's docs, and the product path refuses to compute without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for header in ("descent_cuda.h", "descent_api.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(dsc_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(built_library):
    lib = ctypes.CDLL(built_library.LIB_PATH)
    names = declared_symbols()
    assert len(names) > 80
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_device_means_error_not_fallback(built_library):
    d = built_library
    if d.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(d.DescentError):
        d.Environment(0)
    env = d.Environment(-1)
    a = env.static_parameter([4], "a")
    b = env.static_parameter([4], "b")
    g = env.build_graph(lambda s: s.write_parameter_value(b, s.parameter_value(a) + 1.0))
    with pytest.raises(d.DescentError, match="no CPU execution path"):
        env.run(g, 0)
    with pytest.raises(d.DescentError, match="no CPU execution path"):
        env.read_parameter_to_vec(b)


def test_product_package_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "descent_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".h")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(root, f)


@pytest.mark.parametrize("network,m", [("linear", 1000), ("single-layer-dropout", 1000), ("conv-net", 1000), ("conv-blur-net", 100),
                                       ("siren", 16384), ("relu-pe", 16384), ("multi-hash", 16384)])
def test_generated_kernels_compile_for_sm100a(built_library, host_env, network, m):
    ex = host_env.example(network, m)
    source = ex.train_graph.kernel_source()
    assert "tcgen05" not in source  # JIT clusters are SIMT; the tcgen05 GEMM is the precompiled dsc_gemm_tf32
    assert built_library.nvrtc_compile(source) > 0


@pytest.mark.parametrize("network,m,kwargs", [("conv-net", 8192, {}), ("conv-blur-net", 256, {}), ("multi-hash", 262144, {}), ("single-layer", 1024, {}),
                                              ("sentiment", 64, {"image_width": 512, "image_height": 8})])
def test_tensor_core_kernels_compile_for_sm100a(built_library, host_env, network, m, kwargs):
    """The TF32 variant of the step: halo-tiled conv kernels, gathered tcgen05 GEMMs, epilogues, column sums (all JIT
    templates with inline tcgen05 / TMEM PTX) must at least compile without a GPU."""
    ex = host_env.example(network, m, **kwargs)
    source = ex.train_graph.kernel_source(tf32=True)
    assert built_library.nvrtc_compile(source) > 0


def test_sentiment_example_structure(built_library, host_env):
    """examples/sentiment/main.rs:119-170: embedding [vocab, 128], LSTM 128 -> 64 (4 gates x (wi, wh, b)), dense 64 -> 3."""
    vocab, words = 500, 32
    ex = host_env.example("sentiment", 16, image_width=vocab, image_height=words)
    shapes = {p.name(): p.shape() for p in ex.parameters}
    assert shapes["em"] == (vocab, 128) and shapes["w"] == (64, 3) and shapes["forget_wh"] == (64, 64) and shapes["cell_wi"] == (128, 64)
    assert sum(p.element_count() for p in ex.parameters) == vocab * 128 + 4 * (128 * 64 + 64 * 64 + 64) + 64 * 3 + 3
    assert ex.x.shape() == (16, words, 1)
    small = host_env.example("sentiment", 16, image_width=64, image_height=4)  # strict-FP32 kernels of a 4-word unroll
    assert built_library.nvrtc_compile(small.train_graph.kernel_source()) > 0


def test_parameter_counts_match_reference_readme(host_env):
    """examples/image_fit/README.md:31-34 and the conv-net size quoted in SURVEY.md section 5."""
    expected = {"relu": 44099, "relu-pe": 51779, "siren": 44099, "multi-hash": 43977, "conv-net": 204618, "linear": 7850}
    for network, count in expected.items():
        ex = host_env.example(network, 16)
        assert sum(p.element_count() for p in ex.parameters) == count, network


def test_hash_grid_level_sizes(host_env):
    """grid sizes 2,3,6,12,23,43,80,149,276,512 -> table rows."""
    ex = host_env.example("multi-hash", 16)
    rows = [p.shape()[0] for p in ex.parameters if p.name() == "t"]
    assert rows == [9, 16, 49, 169, 576, 1936, 4096, 4096, 4096, 4096]


def test_array_api_graph_structure(host_env):
    """docs/array_api_values.svg: one MatMul cluster + one per-element cluster {Mul, Mul, Add, Add};
    docs/array_api_grad.svg: a single 7-op per-element cluster {Sin, Cos, Add, Add, Mul, Mul, Sub}."""
    env = host_env
    m = env.static_parameter([3, 3], "m")
    x = env.static_parameter([3, 1], "x")
    y = env.static_parameter([3, 1], "y")
    z = env.static_parameter([3, 1], "z")
    scope = env.scope()
    scope.write_parameter_value(z, 2.0 * scope.parameter_value(m).matmul(scope.parameter_value(x)) + scope.parameter_value(y) * scope.parameter_value(y) + 1.0)
    g = scope.build_graph().export_json()
    labels = sorted(c["label"].split(" ")[0] for c in g["clusters"])
    assert labels == ["MatMul", "PerElement"]
    ops = sorted(n["kind"] for n in g["nodes"] if n["op"] == "Binary")
    assert ops == ["Add", "Add", "Mul", "Mul"]

    xp = env.trainable_parameter([1], "x")
    scope = env.scope()
    xd = scope.parameter(xp)
    yd = xd.sin()
    (yd.square() + yd * 3.0).set_loss()
    scope.write_parameter_value(xp, xd.value() - 0.1 * xd.loss_grad())
    g = scope.build_graph().export_json()
    assert len(g["clusters"]) == 1
    kinds = sorted(n["kind"] for n in g["nodes"] if n["op"] in ("Unary", "Binary"))
    assert kinds == ["Add", "Add", "Cos", "Mul", "Mul", "Sin", "Sub"]  # the *(1/m) with m = 1 was simplified away


def test_conv_im2col_is_not_materialised(host_env):
    """The reference copies the [M, g, K] window matrix (graph.rs:262-284); here the MatMul reads the layer
    input directly through a view chain (zero copies: BASELINE.json north_star (a))."""
    ex = host_env.example("conv-net", 8)
    g = ex.train_graph.export_json()
    by_id = {n["id"]: n for n in g["nodes"]}
    convs = [n for n in g["nodes"] if n["op"] == "MatMul" and n["mode"] == "Rows"]
    assert len(convs) == 2
    for n in convs:
        src = by_id[n["args"][0]["src"]]
        assert len(n["args"][0]["chain"]["views"]) >= 1          # the 7-D window view (+ a group permute when g > 1)
        assert np.prod(src["shape"]) < n["args"][0]["chain"]["output_count"]  # reads the un-expanded image
    assert not any(c["label"].startswith("PerElement (1 ops)") for c in g["clusters"])  # no copy kernels at all


def test_split_k_reduce_is_absorbed(host_env):
    """MatMul [r, ...] + Reduce(axis 0) (array.rs:515) becomes one GEMM cluster; K = 1568 -> r = 2 in the graph."""
    ex = host_env.example("conv-net", 8)
    g = ex.train_graph.export_json()
    mm = [n for n in g["nodes"] if n["op"] == "MatMul" and n["shape"][0] > 1]
    assert mm
    by_cluster = {}
    for n in g["nodes"]:
        by_cluster.setdefault(n["cluster"], []).append(n["op"])
    for n in mm:
        ops = sorted(by_cluster[n["cluster"]])
        # the split-K Reduce always; further Reduce nodes are the bias-gradient chain that shares the GEMM's B operand
        assert ops[:2] == ["MatMul", "Reduce"] and set(ops[2:]) <= {"Reduce"}, ops


def test_conv_net_fusions_are_in_place(host_env):
    """The cluster-level rewrites DESIGN.md section 1 lists, on conv-net's training step: conv backward-input absorbs
    col2im and both Unpads, bias + activation run in the conv GEMMs' epilogues, the bias-gradient Reduce chains ride on
    the weight-gradient GEMMs, pool backward and activation backward are one per-element kernel, and the
    pre-activations are never stored (no cluster writes a tensor that only a `x > 0` test reads)."""
    m = 4096  # the column sums join a GEMM only when its reduction is long enough to be worth streaming once (K >= 4096)
    ex = host_env.example("conv-net", m)
    labels = [c["label"] for c in ex.train_graph.export_json()["clusters"]]
    assert sum("MatMul+WindowsToImage" in l and l.endswith("+Unpad+Unpad") for l in labels) == 1, labels
    assert sum(l.startswith("MatMul") and " + PerElement" in l for l in labels) >= 3, labels          # conv1, conv2, dense layers
    assert sum("+ColumnSum" in l for l in labels) == 4, labels                                          # two convs, two dense layers
    assert not any(l.startswith("WindowsToImage") or l.startswith("Unpad") for l in labels), labels
    big = [l for l in labels if l.startswith("PerElement") and "[%d]" % (m * 28 * 28 * 16) in l]
    assert len(big) == 1, big  # forward bias+leaky lives in the conv epilogue; backward pool+leaky is the one kernel left
    assert len(labels) <= 45, len(labels)


def test_invalid_shapes_are_errors_not_crashes(host_env):
    """The reference asserts / panics on these (shape.rs:44-48: every extent > 0 and at most 7 axes; reshape and broadcast
    mismatches, shape.rs:421-455,578-602); across the C ABI they surface as error codes (DescentError here)."""
    d = __import__("descent_b200")
    for shape in ([0, 3], [4, -1], [2] * 8):
        with pytest.raises(d.DescentError):
            host_env.static_parameter(shape, "bad")
    a = host_env.static_parameter([4, 3], "a")
    b = host_env.static_parameter([5, 2], "b")
    scope = host_env.scope()
    with pytest.raises(d.DescentError):
        scope.parameter_value(a).reshape([5, 3])
    with pytest.raises(d.DescentError):
        scope.parameter_value(a) + scope.parameter_value(b)
    with pytest.raises(d.DescentError):
        host_env.run(scope.build_graph(), 0)  # a host-only environment has no execution path at all


def test_conv_backward_producers_are_evaluated_inside_the_gemm_loaders(built_library, host_env):
    """Operand prologues (graph.hpp OperandPrologue): max-pool backward o leaky-relu backward, the per-element kernel
    that writes dY of the first convolution (reference: select_eq + leaky backward clusters, array.rs:1049-1059), has
    one consumer -- the weight-gradient GEMM with its bias column sums -- and is evaluated inside that kernel's operand
    loader: the [m, 28, 28, 16] gradient array is never written.  The second convolution's dY feeds two GEMMs (weight
    gradient and backward-input); evaluating it twice measured slower, so it stays a kernel (graph.cpp)."""
    ex = host_env.example("conv-net", 1024)
    for tf32 in (True, False):
        source = ex.train_graph.kernel_source(tf32=tf32)
        assert source.count("computed while loading: PerElement (9 ops) [12845056]") == 1   # conv1 dY -> thin weight-gradient kernel
        assert "\n// PerElement (9 ops) [12845056]\n" not in source
        assert "\n// PerElement (9 ops) [6422528]\n" in source and "computed while loading: PerElement (9 ops) [6422528]" not in source
        assert built_library.nvrtc_compile(source) > 0


def test_round2_fusions_in_the_conv_net_step(host_env):
    """Cluster-level rewrites of round 2 on conv-net's training step: softmax cross-entropy is ONE row
    kernel, every parameter update and running sum is ONE grouped launch on the last level, both max-pool forwards ride in
    their convolution's cluster, and the step has at most 25 clusters (34 in round 1)."""
    ex = host_env.example("conv-net", 1000)
    clusters = ex.train_graph.export_json()["clusters"]
    labels = [c["label"] for c in clusters]
    rows = [l for l in labels if l.startswith("Row (")]
    assert len(rows) == 1 and "[1000, 10]" in rows[0], labels
    assert not any(l.startswith("Reduce (k=10)") for l in labels), labels           # the class-axis reductions are inside the row kernel
    assert sum("+MaxPool" in l for l in labels) == 2 and not any(l.startswith("Reduce (k=4)") for l in labels), labels
    last = max(c["level"] for c in clusters)
    tail = [c for c in clusters if c["level"] == last]
    assert len(tail) == 1 and tail[0]["label"].startswith("PerElementGroup (10 programs)"), [c["label"] for c in tail]  # 8 tensors + loss + accuracy
    assert len(clusters) <= 25, len(clusters)
    # the same pass on multi-hash: Adam's 16 tensors and the loss sum in one launch (the four 4096-row tables share a program)
    mh = host_env.example("multi-hash", 4096).train_graph.export_json()["clusters"]
    last = max(c["level"] for c in mh)
    assert [c["label"].split(" [")[0] for c in mh if c["level"] == last] == ["PerElementGroup (14 programs)"]


def test_kernels_refuse_arrays_beyond_32_bit_indexing(built_library, host_env):
    """Generated kernels index with 32-bit integers: a per-GPU mini-batch whose activations exceed 2^31 elements
    is refused with a message instead of wrapping."""
    ex = host_env.example("conv-net", 200000)  # [m, 28, 28, 16] = 2.5e9 elements
    with pytest.raises(built_library.DescentError, match="index with 32 bits"):
        ex.train_graph.kernel_source()


def test_round2_fusions_are_planned(built_library, host_env):
    """Structure of the image_fit multi-hash step after the round-2 passes (no GPU): one dense chain over the MLP's twelve
    clusters, emitted as ONE tcgen05 kernel under TF32 only; the nine concat selects folded into one kernel; the ten
    tables' scatter_adds grouped into one partial + one sum launch with their values computed in the loader."""
    ex = host_env.example("multi-hash", 262144, image_width=1024, image_height=1024)
    graph = ex.train_graph.export_json()
    chains = graph["dense_chains"]
    assert len(chains) == 1 and chains[0]["widths"] == [20, 64, 64, 3] and chains[0]["rows"] == 262144
    assert len(chains[0]["clusters"]) == 12 and len(chains[0]["forward"]) == 3 and all(b >= 0 for b in chains[0]["backward"])
    labels = [c["label"] for c in graph["clusters"]]
    assert sum(l.startswith("ScatterAdd") for l in labels) == 10
    assert sum(l.startswith("PerElement (5 ops)") for l in labels) <= 1  # the concat chain is one select kernel, not nine
    tf32, strict = ex.train_graph.kernel_source(tf32=True), ex.train_graph.kernel_source(tf32=False)
    assert "dense chain: 3 layers, widths 20 64 64 3" in tf32 and "tcgen05.mma.cta_group::1.kind::tf32" in tf32
    assert "dense chain" not in strict  # strict FP32: the chain's clusters run one by one
    for source in (tf32, strict):
        assert source.count("_parts(") == 1 and source.count("_sums(") >= 1  # grouped scatter launches
        assert "_value0(" in source  # scatter values evaluated while loading
    assert tf32.count("__global__") <= 20, tf32.count("__global__")
    # wide heads are detected but not fused (shared memory), networks without an MLP head have no chain
    assert host_env.example("relu-pe", 65536, image_width=512, image_height=512).train_graph.export_json()["dense_chains"][0]["widths"] == [32, 256, 128, 64, 32, 3]
    assert "dense chain" not in host_env.example("relu-pe", 65536, image_width=512, image_height=512).train_graph.kernel_source(tf32=True)
    assert host_env.example("conv-net", 64).train_graph.export_json()["dense_chains"] == []


def test_conv_blur_net_kernels(built_library, host_env):
    """MaxBlurPool2D (module.rs:139-163, 221-245): depthwise blur forward as per-output dot products, its strided backward
    as one gather cluster (MatMul + WindowsToImage + two Unpads), the overlapping max-pool backward inside the col2im gather."""
    ex = host_env.example("conv-blur-net", 64)
    labels = [c["label"] for c in ex.train_graph.export_json()["clusters"]]
    assert sum("MatMul+WindowsToImage/2x2" in l and l.endswith("+Unpad+Unpad") for l in labels) == 2, labels
    assert not any(l.startswith("Unpad") for l in labels)
    source = ex.train_graph.kernel_source(tf32=False)
    assert source.count("[gather: one thread per image-gradient element]") == 2
    assert "_opA1(" in source  # window values of the stand-alone col2im come from an operand prologue
