"""Data-parallel semantics on CPU: rank graphs built for world = 2 (1/m_global loss-gradient
seed, one all-reduce bucket before weight decay and Adam, rank-offset dropout indices) reproduce the single-rank
step on the full batch.  The lock-step oracle sums AllReduce nodes; a second test runs the same reduction through
torch.distributed (gloo, world_size 2) the way bench.py's ranks would."""
import os

import numpy as np
import pytest

from helpers import init_example_params, max_rel_err, synthetic_batch
from oracle import interp, run_graph


def build(d, world, rank, network, m_local):
    env = d.Environment(-1)
    if world > 1:
        env.set_data_parallel_for_tracing(world, rank)
    return env, env.example(network, m_local)


@pytest.mark.parametrize("network", ["single-layer-dropout", "conv-net", "multi-hash"])
def test_two_rank_step_equals_single_rank_step(built_library, network):
    d = built_library
    m = 8
    env1, ex1 = build(d, 1, 0, network, m)
    rng = np.random.default_rng(11)
    params = init_example_params(ex1, rng)
    x, y = synthetic_batch(ex1, rng)
    params[ex1.x.id], params[ex1.y.id] = x, y
    seed = 99
    want = run_graph(ex1.train_graph_json, params, seed)

    graphs, per_rank = [], []
    for rank in range(2):
        env, ex = build(d, 2, rank, network, m // 2)
        assert [p.id for p in ex.parameters] == [p.id for p in ex1.parameters]
        pr = dict(params)
        pr[ex.x.id], pr[ex.y.id] = x[rank * m // 2:(rank + 1) * m // 2], y[rank * m // 2:(rank + 1) * m // 2]
        graphs.append(ex.train_graph_json)
        per_rank.append(pr)
        g = ex.train_graph.export_json()
        ar = [c for c in g["clusters"] if c["label"].startswith("AllReduce")]
        # at most two buckets = two levels: gradients ready by the time the largest one is (reduced on a side stream under
        # the rest of the backward pass) and the late remainder; every reader of a reduced gradient sits behind both
        ar_levels = sorted({c["level"] for c in ar})
        assert len(ar) == len(ex.parameters) and len(ar_levels) <= 2, ar_levels
        ar_ids = {c["outputs"][0] for c in ar}
        for c in g["clusters"]:
            if not c["label"].startswith("AllReduce") and any(i in ar_ids for i in c["inputs"]):
                assert c["level"] > ar_levels[-1], (c["label"], c["level"], ar_levels)
    got = interp.run_graph_data_parallel(graphs, per_rank, seed)
    for p in ex1.parameters + ex1.optimizer_state:
        # relative to each tensor's scale: the sharded sum only reorders f32 additions; Adam's first step
        # (alpha * sign-like ratio) amplifies that on near-zero gradient entries (see tests/test_gpu_networks.py)
        assert max_rel_err(got[0][p.id], want[p.id]) <= (2e-4 if p in ex1.parameters else 1e-5), p.name()
        np.testing.assert_array_equal(got[0][p.id], got[1][p.id])  # replicas stay identical
    total_loss = got[0][ex1.loss_sum.id] + got[1][ex1.loss_sum.id]  # per-rank partial sums, added at read-back
    np.testing.assert_allclose(total_loss, want[ex1.loss_sum.id], rtol=1e-5)


def _gloo_worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import descent_b200 as d
    m = 8
    env, ex = build(d, world, rank, "single-layer-dropout", m // world)
    rng = np.random.default_rng(11)
    env1, ex1 = build(d, 1, 0, "single-layer-dropout", m)
    params = init_example_params(ex1, rng)
    x, y = synthetic_batch(ex1, rng)
    params[ex.x.id], params[ex.y.id] = x[rank * m // world:(rank + 1) * m // world], y[rank * m // world:(rank + 1) * m // world]

    class Rank(interp._Interp):
        pass
    it = interp._Interp(ex.train_graph_json, params, 99, rank)
    live = interp._live_nodes(ex.train_graph_json)
    bucket = []
    for node in it.nodes:
        if node["id"] not in live:
            continue
        if node["op"] == "AllReduce":
            t = torch.from_numpy(it.arg(node, 0).astype(np.float64))
            dist.all_reduce(t)  # gloo sum across the two processes
            it.values[node["id"]] = t.numpy().astype(np.float32)
            bucket.append(node["id"])
            continue
        v = it.eval(node)
        if v is not None:
            it.values[node["id"]] = v
    np.save(os.path.join(tmpdir, "rank%d.npy" % rank), it.outputs[ex.parameters[0].id])
    dist.destroy_process_group()


def test_gloo_all_reduce_world_size_two(built_library, tmp_path):
    import torch.multiprocessing as mp
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(a, b)
    d = built_library
    env1, ex1 = build(d, 1, 0, "single-layer-dropout", 8)
    rng = np.random.default_rng(11)
    params = init_example_params(ex1, rng)
    params[ex1.x.id], params[ex1.y.id] = synthetic_batch(ex1, rng)
    want = run_graph(ex1.train_graph_json, params, 99)
    assert max_rel_err(a, want[ex1.parameters[0].id]) <= 2e-4
