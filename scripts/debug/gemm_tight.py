"""Debug: dense tcgen05 GEMM against the product of TF32-truncated operands, tight tolerance (run under gpurun)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import descent_b200 as d
from oracle.interp import tf32_operand
from test_gpu_gemm_tf32 import device_gemm

env = d.Environment(0)
cases = [(1000, 128, 1568, 1, 1, 6), (1000, 128, 1568, 1, 1, 1), (1568, 128, 1000, 0, 1, 4), (1568, 128, 1000, 0, 1, 1), (1568, 128, 1024, 0, 1, 4),
         (1000, 1568, 128, 1, 0, 1), (1568, 128, 8192, 0, 1, 11), (1000, 304, 784, 0, 1, 1), (1000, 304, 784, 1, 0, 1), (1000, 304, 776, 0, 0, 1), (1000, 304, 776, 1, 1, 1)]
for m, n, k, a_is_mk, b_is_kn, splits in cases:
    rng = np.random.default_rng(m + n + k)
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((k, n)).astype(np.float32)
    a_store = np.ascontiguousarray(a if a_is_mk else a.T)
    b_store = np.ascontiguousarray(b if b_is_kn else b.T)
    got = device_gemm(d, env, a_store, b_store, m, n, k, a_is_mk, b_is_kn, splits)
    if splits > 1:
        got = got.astype(np.float64).sum(0)
    exact = tf32_operand(a, "trunc").astype(np.float64) @ tf32_operand(b, "trunc").astype(np.float64)
    err = np.abs(got - exact)
    rows = np.where(err.max(1) > 1e-4 * np.abs(exact).max())[0]
    print("m=%d n=%d k=%d a_mk=%d b_kn=%d splits=%d: max err / max %.3g; bad rows %d (%s)" % (m, n, k, a_is_mk, b_is_kn, splits, err.max() / np.abs(exact).max(), len(rows), rows[:8]))
    if len(rows):
        # which k range explains the error?  compare with the product that leaves out the last partial k block
        kk = (k // 32) * 32
        part = tf32_operand(a[:, :kk], "trunc").astype(np.float64) @ tf32_operand(b[:kk], "trunc").astype(np.float64)
        print("   without the last partial k block: max err %.3g" % (np.abs(got - part).max() / np.abs(exact).max()))
env.close()
