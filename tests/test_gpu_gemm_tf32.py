"""dsc_gemm_tf32 (TMA + tcgen05 + TMEM) against an exact float64 product, all four operand layouts.
Tolerance: TF32 keeps 10 explicit mantissa bits, so every product carries at most 2^-10 relative error
(2^-11 per operand); FP32 accumulation adds ~K*2^-24.  |err| <= 2^-10 * (|A|.|B|) elementwise."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def device_gemm(d, env, a_store, b_store, m, n, k, a_is_mk, b_is_kn, splits=1):
    lib, ctx = d.lib, env.ctx()
    bufs = []
    for arr in (a_store, b_store, np.zeros((splits, m, n), np.float32)):
        h = ctypes.c_uint64(0)
        assert lib.dsc_alloc(ctx, ctypes.c_size_t(arr.nbytes), ctypes.byref(h)) == 0
        assert lib.dsc_upload(ctx, h, ctypes.c_size_t(0), arr.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(arr.nbytes), ctypes.c_size_t(0), 0) == 0
        bufs.append(h)
    if splits == 1:
        rc = lib.dsc_gemm_tf32(ctx, bufs[0], bufs[1], bufs[2], ctypes.c_int64(m), ctypes.c_int64(n), ctypes.c_int64(k), a_is_mk, b_is_kn)
    else:
        rc = lib.dsc_gemm_tf32_split_k(ctx, bufs[0], bufs[1], bufs[2], ctypes.c_int64(m), ctypes.c_int64(n), ctypes.c_int64(k), a_is_mk, b_is_kn, splits)
    assert rc == 0, lib.dsc_last_error().decode()
    out = np.empty((splits, m, n) if splits > 1 else (m, n), np.float32)
    assert lib.dsc_download(ctx, bufs[2], ctypes.c_size_t(0), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(out.nbytes)) == 0, lib.dsc_last_error().decode()
    for h in bufs:
        lib.dsc_free(ctx, h)
    return out


@pytest.mark.parametrize("a_is_mk", [1, 0])
@pytest.mark.parametrize("b_is_kn", [0, 1])
@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (256, 384, 160), (1024, 256, 4096), (1000, 300, 784), (1568, 128, 8192), (36, 12, 20), (19072, 1024, 96)])
def test_gemm_tf32_layouts(built_library, env, m, n, k, a_is_mk, b_is_kn):
    rng = np.random.default_rng(m + 3 * n + 7 * k + 2 * a_is_mk + b_is_kn)
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((k, n)).astype(np.float32)
    a_store = np.ascontiguousarray(a if a_is_mk else a.T)
    b_store = np.ascontiguousarray(b if b_is_kn else b.T)
    got = device_gemm(built_library, env, a_store, b_store, m, n, k, a_is_mk, b_is_kn)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    bound = 2.0 ** -10 * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)) + 1e-6
    err = np.abs(got - exact)
    assert (err <= bound).all(), "max err/bound %.3g at %s" % ((err / bound).max(), np.unravel_index((err / bound).argmax(), err.shape))


@pytest.mark.parametrize("m,n,k,splits,a_is_mk,b_is_kn", [(1568, 128, 8192, 11, 0, 1), (128, 10 * 4, 8192, 32, 0, 1), (300, 200, 1000, 4, 1, 0)])
def test_gemm_tf32_split_k(built_library, env, m, n, k, splits, a_is_mk, b_is_kn):
    """Slice s of dsc_gemm_tf32_split_k is the product over its own k blocks: each partial is checked on its own, and
    the partials in slice order add up to the full product (the dense-layer weight gradients take this path)."""
    rng = np.random.default_rng(k + splits)
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((k, n)).astype(np.float32)
    a_store = np.ascontiguousarray(a if a_is_mk else a.T)
    b_store = np.ascontiguousarray(b if b_is_kn else b.T)
    got = device_gemm(built_library, env, a_store, b_store, m, n, k, a_is_mk, b_is_kn, splits)
    per = -(-(-(-k // 32)) // splits) * 32  # k elements per slice: ceil(ceil(k / 32) / splits) blocks of 32
    for s in range(splits):
        lo, hi = s * per, min(k, (s + 1) * per)
        exact = a[:, lo:hi].astype(np.float64) @ b[lo:hi].astype(np.float64)
        bound = 2.0 ** -10 * (np.abs(a[:, lo:hi]).astype(np.float64) @ np.abs(b[lo:hi]).astype(np.float64)) + 1e-6
        assert (np.abs(got[s] - exact) <= bound).all(), s
    lib, ctx = built_library.lib, env.ctx()
    rc = lib.dsc_gemm_tf32_split_k(ctx, ctypes.c_uint64(256), ctypes.c_uint64(256), ctypes.c_uint64(256), ctypes.c_int64(128), ctypes.c_int64(128), ctypes.c_int64(96), 1, 1, 4)
    assert rc == 5  # 3 k blocks cannot feed 4 slices


def test_gemm_tf32_rejects_unaligned(built_library, env):
    lib, ctx = built_library.lib, env.ctx()
    rc = lib.dsc_gemm_tf32(ctx, ctypes.c_uint64(256), ctypes.c_uint64(256), ctypes.c_uint64(256), ctypes.c_int64(100), ctypes.c_int64(128), ctypes.c_int64(30), 1, 1)
    assert rc == 5  # DSC_ERR_UNSUPPORTED: the caller falls back to the JIT SIMT GEMM


def test_tf32_operand_rounding_mode(built_library, env, capsys):
    """Which FP32 -> TF32 conversion does tcgen05 kind::tf32 apply to its shared-memory operands?  The oracle's TF32
    emulation (oracle.interp.tf32_operand) must use the same one; asserted: truncation matches to FP32 accumulation noise."""
    from oracle.interp import tf32_operand
    m, n, k = 256, 256, 512
    rng = np.random.default_rng(3)
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((k, n)).astype(np.float32)
    got = device_gemm(built_library, env, a, np.ascontiguousarray(b.T), m, n, k, 1, 0)
    errs = {}
    for mode in ("trunc", "rna"):
        exact = tf32_operand(a, mode).astype(np.float64) @ tf32_operand(b, mode).astype(np.float64)
        errs[mode] = float(np.abs(got - exact).max() / np.abs(exact).max())
    with capsys.disabled():
        print("\ntf32 operand conversion: max rel err vs truncation %.3g, vs round-to-nearest-away %.3g" % (errs["trunc"], errs["rna"]))
    assert min(errs.values()) < 5e-6, errs
    assert errs["trunc"] < errs["rna"], errs
