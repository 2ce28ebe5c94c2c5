"""fashion_mnist on the CUDA backend: the training loop of the reference's examples/fashion_mnist/main.rs:325-440.

    python examples/fashion_mnist.py [single-layer|linear|conv-net|conv-blur-net] [-o adam|descent] [-m 1000] [-e 40] [-t 1]
                                     [--data data/fashion_mnist] [--csv stats.csv] [--tf32] [--show-timings] [--quiet]

Same flow as the reference: reset every trainable parameter from ChaCha20Rng::seed_from_u64(trial), per epoch set the
learning-rate scale 0.5^(epoch/40), shuffle the training indices with that generator, run one graph step per mini-batch
with rand_seed = rng.next_u32(), then evaluate the test set, print and append one CSV row.  The data are the gzip IDX files
of the dataset (`--data`, loaded by the library's front end); without them (this repository cannot download anything) a
synthetic set of the same format is generated in memory so the loop can be exercised."""
import argparse
import gzip
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import descent_b200 as d  # noqa: E402


def synthetic_idx(images, seed):
    """A stand-in with the dataset's format: ten noisy class templates, 28 x 28 bytes, labels 0..9."""
    rng = np.random.default_rng(seed)
    templates = rng.integers(0, 256, (10, 28, 28)).astype(np.float32)
    labels = rng.integers(0, 10, images).astype(np.uint8)
    pixels = np.clip(templates[labels] * 0.6 + rng.normal(0, 40, (images, 28, 28)), 0, 255).astype(np.uint8)
    return (gzip.compress(struct.pack(">IIII", 2051, images, 28, 28) + pixels.tobytes(), 1),
            gzip.compress(struct.pack(">II", 2049, images) + labels.tobytes(), 1))


def load(data_dir, stem, fallback):
    path = os.path.join(data_dir, stem)
    if os.path.exists(path):
        return d.load_gz_bytes(path)  # main.rs:13-19
    return d.gunzip(fallback)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("network", nargs="?", default="single-layer", choices=["linear", "single-layer", "conv-net", "conv-blur-net"])
    ap.add_argument("-o", "--optimizer", default="adam", choices=["adam", "descent"])
    ap.add_argument("-w", "--weight-decay", type=float, default=1.0e-8)
    ap.add_argument("-m", "--mini-batch-size", type=int, default=1000)
    ap.add_argument("-e", "--epoch-count", type=int, default=40)
    ap.add_argument("-t", "--trial-count", type=int, default=1)
    ap.add_argument("--data", default="data/fashion_mnist")
    ap.add_argument("--synthetic-images", type=int, default=10000, help="training images of the stand-in set when --data has no files")
    ap.add_argument("--csv", default="")
    ap.add_argument("--tf32", action="store_true", help="tensor-core GEMMs with TF32 operands (default: strict FP32)")
    ap.add_argument("--show-timings", action="store_true")
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args()
    m = args.mini_batch_size

    env = d.Environment(0)
    env.set_tf32(args.tf32)
    ex = env.example(args.network, m, optimizer=args.optimizer, weight_decay=args.weight_decay)
    fake_train, fake_train_labels = synthetic_idx(args.synthetic_images, 1)
    fake_test, fake_test_labels = synthetic_idx(max(m, args.synthetic_images // 5 // m * m), 2)
    train_images = load(args.data, "train-images-idx3-ubyte.gz", fake_train)
    train_labels = load(args.data, "train-labels-idx1-ubyte.gz", fake_train_labels)
    test_images = load(args.data, "t10k-images-idx3-ubyte.gz", fake_test)
    test_labels = load(args.data, "t10k-labels-idx1-ubyte.gz", fake_test_labels)
    train_count, rows, cols = d.read_images_info(train_images)
    test_count, _, _ = d.read_images_info(test_images)
    assert train_count == d.read_labels_info(train_labels) and test_count == d.read_labels_info(test_labels)
    assert train_count % m == 0 and test_count % m == 0 and rows == 28 and cols == 28  # main.rs:340-355
    stats = open(args.csv, "w") if args.csv else None
    if args.show_timings:
        env.set_options(use_cuda_graph=False, profile_runs=True)

    for trial in range(args.trial_count):
        rng = d.ChaCha20Rng(trial)  # main.rs:362
        for p in ex.parameters:
            env.reset_parameter_rng(p, rng)
        for p in ex.optimizer_state:  # Optimizer::reset_state
            env.zero_fill(p)
        for epoch in range(args.epoch_count):
            env.write(ex.learning_rate_scale, np.array([0.5 ** (epoch / 40.0)], np.float32))  # main.rs:371-375
            env.zero_fill(ex.loss_sum)
            env.zero_fill(ex.accuracy_sum)
            indices = rng.shuffle(np.arange(train_count))  # main.rs:380-382
            for start in range(0, train_count, m):
                batch = indices[start:start + m]
                env.write(ex.x, d.unpack_images(train_images, batch))
                env.write(ex.y, d.unpack_labels(train_labels, batch))
                env.run(ex.train_graph, rng.next_u32())
            if args.show_timings and epoch < 2:
                env.print_timings("training")
            train_loss = env.read_parameter_scalar(ex.loss_sum) / train_count
            train_accuracy = env.read_parameter_scalar(ex.accuracy_sum) / train_count
            env.zero_fill(ex.loss_sum)
            env.zero_fill(ex.accuracy_sum)
            for start in range(0, test_count, m):
                batch = np.arange(start, start + m)
                env.write(ex.x, d.unpack_images(test_images, batch))
                env.write(ex.y, d.unpack_labels(test_labels, batch))
                env.run(ex.test_graph, rng.next_u32())
            test_loss = env.read_parameter_scalar(ex.loss_sum) / test_count
            test_accuracy = env.read_parameter_scalar(ex.accuracy_sum) / test_count
            if not args.quiet:
                print("epoch: %d, loss: %g/%g, accuracy: %g/%g" % (epoch + 1, train_loss, test_loss, train_accuracy, test_accuracy), flush=True)
            if stats:
                if epoch == 0:
                    stats.write("# epoch, train_loss, test_loss, train_accuracy, test_accuracy\n")
                d.write_csv_row(stats, [epoch + 1, float(train_loss), float(test_loss), float(train_accuracy), float(test_accuracy)])
        if stats:
            stats.write("\n")
    env.close()


if __name__ == "__main__":
    main()
