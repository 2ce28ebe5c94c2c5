"""ctypes front end of oracle/cpu_ref.cpp: the multi-threaded CPU restatement used as the reported CPU baseline
(TEST INFRASTRUCTURE -- see oracle/__init__.py; loaded only by tests/ and bench.py's cpu_baseline / --impl reference).

    build()                      g++ -O2 -> oracle/_build/libcpu_ref.so (called from __graft_entry__.build())
    run_graph(graph, params, seed, threads) -> ({parameter id: new value}, seconds)
    check_graph(graph, params, seed, tf32_nodes) -> {parameter id: new value}: "checker mode" -- float64-accumulated sums
        like oracle.interp (and optional TF32 operand truncation on the listed MatMul nodes), for parity tests at
        mini-batch sizes the numpy interpreter takes minutes for (m = 1000 ... 8192)
"""
import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libcpu_ref.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "cpu_ref.cpp")
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(src):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", src, "-o", LIB_PATH], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.cpu_ref_create.restype = ctypes.c_void_p
        _lib.cpu_ref_create.argtypes = [ctypes.c_char_p]
        _lib.cpu_ref_destroy.argtypes = [ctypes.c_void_p]
        _lib.cpu_ref_set_input.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        _lib.cpu_ref_set_output.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        _lib.cpu_ref_run.restype = ctypes.c_double
        _lib.cpu_ref_run.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int]
        _lib.cpu_ref_error.restype = ctypes.c_char_p
        _lib.cpu_ref_set_checker_mode.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _lib.cpu_ref_set_tie_resolution.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int]
        _lib.cpu_ref_ties_seen.restype = ctypes.c_longlong
        _lib.cpu_ref_ties_seen.argtypes = [ctypes.c_void_p]
    return _lib


def hardware_threads():
    return int(lib().cpu_ref_hardware_threads())


class Program:
    """One exported graph, parsed once; run() evaluates a step on `threads` host threads."""

    def __init__(self, graph):
        self._lib = lib()
        self._h = self._lib.cpu_ref_create(json.dumps(graph).encode())
        if not self._h:
            raise RuntimeError(self._lib.cpu_ref_error().decode())
        self.output_shapes = {n["parameter"]: n["shape"] for n in graph["nodes"] if n["op"] == "Output"}
        self._keep = {}

    def set_checker_mode(self, f64_accumulate=True, tf32_nodes=()):
        ids = np.asarray(sorted(int(i) for i in tf32_nodes), dtype=np.int32)
        self._lib.cpu_ref_set_checker_mode(self._h, int(f64_accumulate), ids.ctypes.data if ids.size else None, int(ids.size))

    def set_tie_resolution(self, select_nodes, margin, resolution):
        ids = np.asarray(sorted(int(i) for i in select_nodes), dtype=np.int32)
        self._lib.cpu_ref_set_tie_resolution(self._h, ids.ctypes.data if ids.size else None, int(ids.size), float(margin), int(resolution))

    def ties_seen(self):
        return int(self._lib.cpu_ref_ties_seen(self._h))

    def run(self, params, seed=0, threads=0):
        threads = threads or hardware_threads()
        for pid, v in params.items():
            a = np.ascontiguousarray(v, dtype=np.float32).reshape(-1)
            self._keep[("in", pid)] = a
            self._lib.cpu_ref_set_input(self._h, int(pid), a.ctypes.data, a.size)
        outputs = {}
        for pid, shape in self.output_shapes.items():
            a = np.empty(int(np.prod(shape)), np.float32)
            outputs[pid] = a
            self._lib.cpu_ref_set_output(self._h, int(pid), a.ctypes.data, a.size)
        seconds = self._lib.cpu_ref_run(self._h, ctypes.c_uint32(int(seed) & 0xFFFFFFFF), int(threads))
        if seconds < 0:
            raise RuntimeError(self._lib.cpu_ref_error().decode())
        return {pid: a.reshape(self.output_shapes[pid]) for pid, a in outputs.items()}, seconds

    def close(self):
        if self._h:
            self._lib.cpu_ref_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def run_graph(graph, params, seed=0, threads=0):
    prog = Program(graph)
    try:
        return prog.run(params, seed, threads)
    finally:
        prog.close()


def check_graph(graph, params, seed=0, tf32_nodes=(), threads=0):
    """The step as oracle.interp.run_graph computes it (float64-accumulated Reduce / MatMul sums rounded once, TF32
    truncation on `tf32_nodes`), on all host threads: the checker for large mini-batches."""
    prog = Program(graph)
    try:
        prog.set_checker_mode(True, tf32_nodes)
        return prog.run(params, seed, threads)[0]
    finally:
        prog.close()


def order_sensitive_selects(graph):
    """Ids of the CompareAndSelect nodes whose compared operands (a, b) depend on a MatMul or a Reduce(Sum): their last
    bits depend on the order of additions, so near-ties may resolve either way in a correct implementation.  Selects on
    exact data (labels, one_hot, the dropout hash) are not in the list."""
    nodes = {n["id"]: n for n in graph["nodes"]}
    memo = {}

    def sensitive(i):
        stack = [i]
        while stack:
            j = stack[-1]
            if j in memo:
                stack.pop()
                continue
            n = nodes[j]
            if n["op"] == "MatMul" or (n["op"] == "Reduce" and n.get("kind") == "Sum"):
                memo[j] = True
                stack.pop()
                continue
            todo = [a["src"] for a in n["args"] if a["src"] not in memo]
            if todo:
                stack.extend(todo)
                continue
            memo[j] = any(memo[a["src"]] for a in n["args"])
            stack.pop()
        return memo[i]

    return [n["id"] for n in graph["nodes"] if n["op"] == "Select" and (sensitive(n["args"][0]["src"]) or sensitive(n["args"][1]["src"]))]


def check_graph_with_tie_band(graph, params, seed=0, tf32_nodes=(), margin=1e-6, threads=0):
    """check_graph plus the band its near-tie selects span: returns (outputs, band, ties) where band[pid] is the largest
    absolute difference, over the step's outputs, between the natural evaluation and the evaluations that resolve every
    near-tie (order_sensitive_selects, relative margin `margin`) as true resp. as false, and `ties` counts those elements."""
    selects = order_sensitive_selects(graph)
    prog = Program(graph)
    try:
        prog.set_checker_mode(True, tf32_nodes)
        natural = prog.run(params, seed, threads)[0]
        natural = {pid: v.copy() for pid, v in natural.items()}
        band = {pid: 0.0 for pid in natural}
        ties = 0
        for resolution in (1, 2):
            prog.set_tie_resolution(selects, margin, resolution)
            forced = prog.run(params, seed, threads)[0]
            ties = max(ties, prog.ties_seen())
            for pid, v in forced.items():
                band[pid] = max(band[pid], float(np.abs(v.astype(np.float64) - natural[pid]).max()))
        return natural, band, ties
    finally:
        prog.close()
