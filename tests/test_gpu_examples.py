"""The example drivers (examples/fashion_mnist.py, examples/image_fit.py: the reference's main.rs training loops over the
library's own RNG and data front ends) run end to end on the GPU and learn."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, *args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)] + list(args), capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def test_fashion_mnist_example_learns_the_stand_in_set(tmp_path):
    csv = tmp_path / "stats.csv"
    out = run("fashion_mnist.py", "single-layer", "-m", "500", "-e", "3", "--synthetic-images", "5000", "--csv", str(csv))
    accuracy = [float(v) for v in re.findall(r"accuracy: ([0-9.e+-]+)/", out)]
    assert len(accuracy) == 3 and accuracy[-1] > 0.9, out  # ten noisy templates: separable
    rows = [l for l in csv.read_text().splitlines() if l and not l.startswith("#")]
    assert len(rows) == 3 and rows[0].startswith("1, ")


def test_image_fit_example_fits_a_synthetic_image(tmp_path):
    prefix = tmp_path / "fit"
    out = run("image_fit.py", "multi-hash", "-e", "6", "-m", "4096", "--synthetic-size", "128", "--image-prefix", str(prefix))
    loss = [float(v) for v in re.findall(r"loss: ([0-9.e+-]+)", out)]
    assert len(loss) == 6 and loss[-1] < 0.5 * loss[0], out
    ppm = (tmp_path / "fit_6.ppm").read_bytes()
    assert ppm.startswith(b"P6\n128 128\n255\n") and len(ppm) == 15 + 128 * 128 * 3
