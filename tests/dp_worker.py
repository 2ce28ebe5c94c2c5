"""Rank body for test_two_gpu_data_parallel_matches_single_gpu (launched by torchrun)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from helpers import init_example_params, max_rel_err, synthetic_batch
    workload = sys.argv[1]
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    optimizer = sys.argv[3] if len(sys.argv) > 3 else "adam"
    tf32 = len(sys.argv) > 4 and sys.argv[4] == "tf32"
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import descent_b200 as d
    # reference run: one GPU, full batch
    ref_env = d.Environment(local)
    ref_env.set_tf32(tf32)
    ref = ref_env.example(workload, m, optimizer=optimizer)
    rng = np.random.default_rng(77)
    params = init_example_params(ref, rng)
    batches = [synthetic_batch(ref, rng) for _ in range(steps)]
    seeds = [int(s) for s in rng.integers(0, 2 ** 32, steps)]
    for pid, v in params.items():
        ref_env.write(ref_env.parameter(pid), v)
    for (x, y), s in zip(batches, seeds):
        ref_env.write(ref.x, x)
        ref_env.write(ref.y, y)
        ref_env.run(ref.train_graph, s)
    compared_ref = ref.parameters if steps > 1 else ref.optimizer_state
    want = [ref_env.read(p) for p in compared_ref]
    want_loss = ref_env.read_parameter_scalar(ref.loss_sum)
    # data-parallel run
    env = d.Environment(local)
    uid = [d.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    env.init_data_parallel(world, rank, uid[0])
    env.set_tf32(tf32)
    ex = env.example(workload, m // world, optimizer=optimizer)
    for p_ref, p in zip(ref.parameters + ref.optimizer_state + [ref.loss_sum, ref.accuracy_sum, ref.learning_rate_scale],
                        ex.parameters + ex.optimizer_state + [ex.loss_sum, ex.accuracy_sum, ex.learning_rate_scale]):
        if p is not None:  # (image_fit examples have no accuracy sum)
            env.write(p, params[p_ref.id])
    lo, hi = rank * m // world, (rank + 1) * m // world
    for (x, y), s in zip(batches, seeds):
        env.write(ex.x, x[lo:hi])
        env.write(ex.y, y[lo:hi])
        env.run(ex.train_graph, s)
    got = [env.read(p) for p in (ex.parameters if steps > 1 else ex.optimizer_state)]
    loss = torch.tensor([env.read_parameter_scalar(ex.loss_sum)], device="cuda")
    dist.all_reduce(loss)
    errs = [max_rel_err(g, w) for g, w in zip(got, want)]
    worst = max(errs)
    ok = worst <= 1e-3 and abs(float(loss.item()) - want_loss) <= 1e-4 * abs(want_loss)
    flags = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    labels = [t["label"] for t in env.profile(ex.train_graph, 0, 1)]  # (every rank: the pass contains the collectives)
    if rank == 0:
        print("deviations %s" % ["%.2g" % e for e in errs])
        print("worst deviation %.3g, loss %g vs %g; %d launches, fused: %s" % (worst, float(loss.item()), want_loss, len(labels),
              [l[:40] for l in labels if "DenseChain" in l or "group" in l or "AllReduce" in l]))
        print("DP_OK" if flags.item() == 1.0 else "DP_MISMATCH")
    env.close()
    ref_env.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
