"""Shared helpers for the parity tests: synthetic inputs per SURVEY.md §8d and oracle comparison."""
import numpy as np

SEED_BASE = 0x5EED5EED


def init_example_params(ex, rng, siren=False):
    """Explicit initial values for every parameter an example graph reads (the harness feeds
    identical arrays to oracle and CUDA path instead of re-deriving them from an unpinned host RNG)."""
    params = {}
    for p in ex.parameters:
        shape = p.shape()
        name = p.name()
        if name == "t":  # hash table: U(+-1e-4) (examples/image_fit/main.rs:141-145)
            v = rng.uniform(-1e-4, 1e-4, shape)
        elif name == "w":
            fan_in = shape[0]
            if siren:
                scale = np.sqrt(6.0 / fan_in) * (30.0 if fan_in == 2 else 1.0)
                v = rng.uniform(-scale, scale, shape)
            else:
                v = rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)  # Initializer::for_relu (parameter.rs:17-20)
        elif name == "em":  # Initializer::RandUniform(1.0) (examples/sentiment/main.rs:131-135)
            v = rng.uniform(-1.0, 1.0, shape)
        elif name == "f":
            v = rng.standard_normal(shape) * np.sqrt(2.0 / (shape[2] * shape[3] * shape[4]))
        else:  # biases
            v = rng.uniform(-1, 1, shape) if siren else np.zeros(shape)
        params[p.id] = v.astype(np.float32)
    for p in ex.optimizer_state:
        params[p.id] = np.zeros(p.shape(), np.float32)
    params[ex.loss_sum.id] = np.zeros(1, np.float32)
    if ex.accuracy_sum is not None:
        params[ex.accuracy_sum.id] = np.zeros(1, np.float32)
    params[ex.learning_rate_scale.id] = np.ones(1, np.float32)
    return params


def synthetic_batch(ex, rng, image_width=512):
    xs, ys = ex.x.shape(), ex.y.shape()
    if len(xs) == 3:  # sentiment: word indices as f32, 0 = padding; labels 0..2 (examples/sentiment/main.rs:176-190)
        vocab = [p for p in ex.parameters if p.name() == "em"][0].shape()[0]
        return rng.integers(0, vocab, xs).astype(np.float32), rng.integers(0, 3, ys).astype(np.float32)
    if len(xs) == 4:  # fashion_mnist: x in [0,1), integer labels as f32 (main.rs:56,69)
        return rng.random(xs, dtype=np.float32), rng.integers(0, 10, ys).astype(np.float32)
    w = image_width  # image_fit: random pixels of a synthetic w x w image (BASELINE.json configs 4 / 5: 512 and 1024)
    px = rng.integers(0, w, (xs[0], 2))
    x = ((px + 0.5) * (2.0 / w) - 1.0).astype(np.float32)  # pixel centres (image_fit/main.rs:379-382)
    return x, rng.random(ys, dtype=np.float32)


def fill_missing_inputs(env, graph_json, params):
    """Parameters the graph reads that the caller did not set (fixed blur filters etc.) are read back from the device."""
    for node in graph_json["nodes"]:
        if node["op"] == "Input" and node["parameter"] not in params:
            params[node["parameter"]] = env.read(env.parameter(node["parameter"]))
    return params


def max_rel_err(got, want):
    got = np.asarray(got, np.float64).reshape(-1)
    want = np.asarray(want, np.float64).reshape(-1)
    scale = max(float(np.abs(want).max()), 1e-30)
    return float(np.abs(got - want).max()) / scale


def upload(env, params):
    for pid, v in params.items():
        env.write(env.parameter(pid), v)
