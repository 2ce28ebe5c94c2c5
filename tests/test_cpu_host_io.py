"""Host side of the examples: the reference examples' random numbers (ChaCha20Rng and the
rand 0.8 sampling rules) and their data front ends (gzip + IDX, baseline JPEG, PPM / CSV output) through the C ABI.
No GPU: these run on the CPU box."""
import gzip
import io
import struct

import numpy as np
import pytest

from oracle import host_rng

# rand_chacha's own test_chacha_true_values_a (all-zero seed) = the zero-key / zero-nonce ChaCha20 keystream
# 76 b8 e0 ad a0 f1 3d 90 40 5d 6a e5 53 86 bd 28 ... read as little-endian words
CHACHA20_ZERO_KEY_BLOCK0 = [0xade0b876, 0x903df1a0, 0xe56a5d40, 0x28bd8653, 0xb819d2bd, 0x1aed8da0, 0xccef36a8, 0xc70d778b,
                            0x7c5941da, 0x8d485751, 0x3fe02477, 0x374ad8b8, 0xf4b8436a, 0x1ca11815, 0x69b687c3, 0x8665eeb2]


def test_oracle_chacha20_block_known_answer():
    assert host_rng.chacha20_block([0] * 8, 0) == CHACHA20_ZERO_KEY_BLOCK0
    rng = host_rng.ChaCha20Rng(bytes(32))
    assert [rng.next_u32() for _ in range(16)] == CHACHA20_ZERO_KEY_BLOCK0


@pytest.mark.parametrize("seed", [0, 1, 4, 0xDEADBEEFCAFE])
def test_library_rng_matches_the_restated_crates(built_library, seed):
    d = built_library
    got, want = d.ChaCha20Rng(seed), host_rng.ChaCha20Rng.seed_from_u64(seed)
    assert [got.next_u32() for _ in range(70)] == [want.next_u32() for _ in range(70)]  # crosses a 64-word refill
    assert [got.next_u64() for _ in range(40)] == [want.next_u64() for _ in range(40)]
    # one word left in the buffer, then next_u64 straddles the refill (BlockRng::next_u64's third branch)
    a, b = d.ChaCha20Rng(seed), host_rng.ChaCha20Rng.seed_from_u64(seed)
    for _ in range(63):
        assert a.next_u32() == b.next_u32()
    assert a.next_u64() == b.next_u64() and a.next_u32() == b.next_u32()
    u = got.open01(1000)
    assert u.dtype == np.float32 and (u > 0).all() and (u < 1).all()
    np.testing.assert_array_equal(u[:50], np.array([want.open01() for _ in range(1000)], np.float32)[:50])
    for bound in (1, 2, 3, 10, 512, 1024, 60000, (1 << 31) + 5):
        g, w = got.gen_range(0, bound, u32=True), want.gen_range(0, bound, u32=True)
        assert g == w and 0 <= g < bound
        g, w = got.gen_range(3, 3 + bound), want.gen_range(3, 3 + bound)
        assert g == w and 3 <= g < 3 + bound
    pairs = got.gen_range_pairs(512, 300, 20)
    assert pairs.tolist() == [[want.gen_range(0, 512), want.gen_range(0, 300)] for _ in range(20)]
    shuffled = got.shuffle(np.arange(1000))
    assert shuffled.tolist() == want.shuffle(range(1000)) and sorted(shuffled.tolist()) == list(range(1000))


def test_gen_range_is_uniform_enough(built_library):
    rng = built_library.ChaCha20Rng(7)
    counts = np.bincount([rng.gen_range(0, 6, u32=True) for _ in range(60000)], minlength=6)
    assert counts.min() > 9500 and counts.max() < 10500


def test_reset_parameter_draws_like_the_reference(host_env, built_library):
    """environment.rs:16-40: RandNormal = scale * sqrt(-2 ln u1) * cos(2 pi u2), RandUniform = scale * (2 u - 1), Open01 draws in order."""
    d = built_library
    # host-only environments cannot hold parameter data: the rule itself is checked on the generator's draws
    rng, ref = d.ChaCha20Rng(3), host_rng.ChaCha20Rng.seed_from_u64(3)
    u = rng.open01(6)
    want = np.array([ref.open01() for _ in range(6)], np.float32)
    np.testing.assert_array_equal(u, want)
    normal = np.float32(0.5) * np.sqrt(np.float32(-2.0) * np.log(u[0::2])) * np.cos(np.float32(2.0 * np.pi) * u[1::2])
    assert np.isfinite(normal).all()


def _idx_images(images):
    n, rows, cols = images.shape
    return struct.pack(">IIII", 2051, n, rows, cols) + images.tobytes()


def test_gzip_idx_front_end(built_library, tmp_path):
    d = built_library
    rng = np.random.default_rng(5)
    images = rng.integers(0, 256, (37, 28, 28), dtype=np.uint8)
    labels = rng.integers(0, 10, 37, dtype=np.uint8)
    ipath, lpath = tmp_path / "images.gz", tmp_path / "labels.gz"
    ipath.write_bytes(gzip.compress(_idx_images(images)))
    lpath.write_bytes(gzip.compress(struct.pack(">II", 2049, 37) + labels.tobytes()))
    ibytes, lbytes = d.load_gz_bytes(str(ipath)), d.load_gz_bytes(str(lpath))
    assert d.read_images_info(ibytes) == (37, 28, 28) and d.read_labels_info(lbytes) == 37
    idx = [36, 0, 5, 5, 17]
    np.testing.assert_array_equal(d.unpack_images(ibytes, idx), images[idx].reshape(5, 784).astype(np.float32) / np.float32(255.0))
    np.testing.assert_array_equal(d.unpack_labels(lbytes, idx), labels[idx].astype(np.float32))
    assert d.gunzip(gzip.compress(b"")) == b""
    with pytest.raises(d.DescentError):
        d.unpack_images(ibytes, [37])
    with pytest.raises(d.DescentError):
        d.read_images_info(lbytes)  # wrong magic
    with pytest.raises(d.DescentError):
        d.gunzip(gzip.compress(b"x" * 1000)[:-12])  # truncated stream
    with pytest.raises(d.DescentError):
        d.load_gz_bytes(str(tmp_path / "missing.gz"))


def _synthetic_image(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(x / 9.0 + seed), 128 + 100 * np.cos(y / 7.0), 128 + 90 * np.sin((x + y) / 13.0)], -1)
    return np.clip(base + rng.normal(0, 6, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("subsampling,size", [(0, (64, 48)), (1, (64, 64)), (1, (70, 45)), (2, (64, 64)), (2, (37, 51)), ("grey", (40, 24))])
def test_baseline_jpeg_decoder_matches_libjpeg(built_library, subsampling, size):
    """Files written by libjpeg (Pillow): 4:4:4, 4:2:2 (the layout of the reference's data/images/cat.jpg), 4:2:0, greyscale,
    ragged sizes.  The decoder restates the IJG integer IDCT, fancy upsampling and colour conversion, so it must agree with
    libjpeg's own decode to the last bit."""
    from PIL import Image
    d = built_library
    w, h = size
    pixels = _synthetic_image(h, w, 3)
    image = Image.fromarray(pixels[..., 0] if subsampling == "grey" else pixels)
    buf = io.BytesIO()
    if subsampling == "grey":
        image.save(buf, "JPEG", quality=90)
    else:
        image.save(buf, "JPEG", quality=90, subsampling=subsampling)
    data = buf.getvalue()
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    got = d.decode_jpeg_rgb(data)
    assert got.shape == want.shape
    diff = np.abs(got.astype(int) - want.astype(int))
    assert diff.max() <= 1, (diff.max(), np.argwhere(diff > 1)[:5])
    assert (diff == 0).mean() > 0.99


def test_jpeg_restart_intervals_and_errors(built_library):
    from PIL import Image
    d = built_library
    pixels = _synthetic_image(48, 80, 1)
    buf = io.BytesIO()
    Image.fromarray(pixels).save(buf, "JPEG", quality=85, subsampling=2, restart_marker_blocks=2)
    want = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
    got = d.decode_jpeg_rgb(buf.getvalue())
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
    buf = io.BytesIO()
    Image.fromarray(pixels).save(buf, "JPEG", progressive=True)
    with pytest.raises(d.DescentError, match="baseline"):
        d.decode_jpeg_rgb(buf.getvalue())
    with pytest.raises(d.DescentError):
        d.decode_jpeg_rgb(b"not a jpeg")


def test_ppm_and_csv_output(built_library, tmp_path):
    d = built_library
    rgb = np.array([[[0.0, 0.5, 1.0], [1.2, -0.1, 0.25]]], np.float32)
    path = tmp_path / "out.ppm"
    d.write_ppm(str(path), rgb)
    raw = path.read_bytes()
    assert raw.startswith(b"P6\n2 1\n255\n")
    assert list(raw[-6:]) == [0, 128, 255, 255, 0, 64]  # x * 255 + 0.5, clamped, truncated (image_fit/main.rs:423-426)
    out = io.StringIO()
    d.write_csv_row(out, ["conv-net", 3, 0.25])
    assert out.getvalue() == '"conv-net", 3, 0.25\n'


def test_op_level_entry_points_plan_without_a_device(host_env):
    """dsc_op_* (include/descent_api.h): every op builds its graph and names its buffers on a host-only environment."""
    ops = {
        "conv fwd": (host_env.op_conv2d(64, 14, 14, 16, 32, 3, 3, pad=1, groups=2), {"x": (64, 14, 14, 16), "filter": (2, 16, 3, 3, 8), "y": (64, 14, 14, 32)}),
        "conv bwd": (host_env.op_conv2d(64, 14, 14, 16, 32, 3, 3, pad=1, groups=2, backward=True), {"dy": (64, 14, 14, 32), "dx": (64, 14, 14, 16), "dfilter": (2, 16, 3, 3, 8)}),
        "scatter": (host_env.op_scatter_add(577, 2, 1000), {"table": (577, 2), "values": (1000, 2), "indices": (1000,)}),
        "xent": (host_env.op_softmax_cross_entropy(100, 10), {"z": (100, 10), "y": (100, 1), "loss": (100, 1), "accuracy": (100, 1), "dz": (100, 10)}),
        "adam": (host_env.op_adam_step([10, 20], 0.01), {"theta0": (10,), "grad1": (20,), "state0": (1,)}),
    }
    for name, (op, buffers) in ops.items():
        for buffer, shape in buffers.items():
            assert tuple(op.parameter(buffer).shape()) == shape, (name, buffer)
    with pytest.raises(Exception):
        ops["xent"][0].parameter("nope")
