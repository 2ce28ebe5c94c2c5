"""image_fit on the CUDA backend: the training loop of the reference's examples/image_fit/main.rs:278-440.

    python examples/image_fit.py [relu|relu-pe|siren|multi-hash] [--image data/images/cat.jpg] [-e 200] [-m 16384]
                                 [--csv stats.csv] [--image-prefix out] [--strict] [--quiet]

Same flow as the reference: load the JPEG (the library's baseline decoder), reset the parameters from
ChaCha20Rng::seed_from_u64(0), per epoch set the learning-rate scale min(t/10, 1) * 0.5^(t/40), draw (width*height)/m
mini-batches of random pixels with rng.gen_range, run one step each with rand_seed = rng.next_u32(), print the loss, and at
the end evaluate the whole image and write it (as PPM; the reference writes a JPEG).  Without an image file a synthetic
one is generated so the loop can be exercised."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import descent_b200 as d  # noqa: E402


def synthetic_image(size):
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    rgb = np.stack([0.5 + 0.5 * np.sin(9 * x + 3 * y), 0.5 + 0.5 * np.cos(7 * y * (1 + x)), (x * y + 0.25 * np.sin(40 * x)) % 1.0], -1)
    return (rgb * 255).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("network", nargs="?", default="multi-hash", choices=["relu", "relu-pe", "siren", "multi-hash"])
    ap.add_argument("--image", default="data/images/cat.jpg")
    ap.add_argument("--synthetic-size", type=int, default=256)
    ap.add_argument("-e", "--epoch-count", type=int, default=200)
    ap.add_argument("-m", "--mini-batch-size", type=int, default=1 << 14)
    ap.add_argument("--csv", default="")
    ap.add_argument("--image-prefix", default="")
    ap.add_argument("--strict", action="store_true", help="strict FP32 GEMMs (default: TF32 tensor-core operands)")
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args()

    data = d.decode_jpeg_rgb(open(args.image, "rb").read()) if os.path.exists(args.image) else synthetic_image(args.synthetic_size)  # main.rs:278-282
    height, width = data.shape[:2]
    m = args.mini_batch_size
    env = d.Environment(0)
    env.set_tf32(not args.strict)
    ex = env.example(args.network, m, image_width=width, image_height=height)
    rng = d.ChaCha20Rng(0)  # main.rs:352
    for p in ex.parameters:
        env.reset_parameter_rng(p, rng)
    for p in ex.optimizer_state:
        env.zero_fill(p)
    stats = open(args.csv, "w") if args.csv else None
    flat = data.reshape(-1, 3)
    loss = float("nan")
    for epoch in range(args.epoch_count):
        t = epoch + 0.5
        lr_scale = min(t / 10.0, 1.0) * 0.5 ** (t / 40.0)  # main.rs:363-364
        env.write(ex.learning_rate_scale, np.array([lr_scale], np.float32))
        env.zero_fill(ex.loss_sum)
        batches = (width * height) // m
        for _ in range(max(1, batches)):
            pixels = rng.gen_range_pairs(width, height, m).astype(np.int64)  # main.rs:376-378: two usize draws per pixel, in this order
            x0, x1 = pixels[:, 0], pixels[:, 1]
            x = np.stack([(x0 + 0.5) * (2.0 / width) - 1.0, (x1 + 0.5) * (2.0 / height) - 1.0], -1).astype(np.float32)
            y = flat[x1 * width + x0].astype(np.float32) / np.float32(255.0)
            env.write(ex.x, x)
            env.write(ex.y, y)
            env.run(ex.train_graph, rng.next_u32())
        loss = env.read_parameter_scalar(ex.loss_sum) / m  # main.rs:402 divides by m, as here
        if not args.quiet:
            print("epoch: %d, lr_scale: %g, loss: %g" % (epoch + 1, lr_scale, loss), flush=True)
        if stats:
            if epoch == 0:
                stats.write("# epoch, loss\n")
            d.write_csv_row(stats, [epoch + 1, float(loss)])
    if args.image_prefix and ex.test_graph is not None:
        env.run(ex.test_graph, rng.next_u32())  # main.rs:419-435
        d.write_ppm("%s_%d.ppm" % (args.image_prefix, args.epoch_count), env.read(ex.image).reshape(height, width, 3))
    env.close()
    return loss


if __name__ == "__main__":
    main()
