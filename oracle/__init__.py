"""CPU oracle for the descent hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker or the reported CPU baseline -- never on the product path
(descent_b200 does not import it and fails loudly without its CUDA library).

Two parts.  `oracle.cpu_ref` (oracle/cpu_ref.cpp, built by `make -C oracle`) is a multi-threaded C++ port of the
same op semantics, laid out the way the reference's kernels compute (one invocation per element, sequential-K
reduce, float-atomic scatter); it is what bench.py times as `cpu_baseline` / `--impl reference`, and it is checked
against the numpy interpreter and the reference's known answers in tests/test_oracle_kat.py.  The parity oracle
proper: a numpy restatement of the *semantics* of the reference (sjb3d/descent) for every op of
its graph IR -- `oracle.interp` interprets the raw op graph the frontend exports as JSON, one function
per `Op` variant, each citing the reference file:line it follows.

Pinning: the reference cannot be built or run in this image (no Rust toolchain, no Vulkan; SURVEY.md
§0), so the oracle is pinned against the known-answer values of the reference's own tests
(src/lib.rs:26-231, examples/array_api/main.rs:24) in tests/test_oracle_kat.py, and against the
derived bit-exact vectors of SURVEY.md Appendix D (pcg / rand / hash-grid indices).  Floating-point
details that no reference test pins (transcendental ulps, FMA contraction, summation order) are
"parity unpinned": the oracle defines them as IEEE f32 per-element ops without contraction and
float64-accumulated sums rounded once, and the parity tests compare within the tolerances
BASELINE.json states (1e-5 relative for strict-FP32 paths).
"""
from .interp import run_graph, run_graph_data_parallel, apply_chain, pcg, rand_from_index  # noqa: F401
