"""Write the generated CUDA translation unit of an example's training graph (no GPU needed).

usage: python scripts/dump_kernels.py conv-net 8192 out.cu [--strict]
Then `nvcc -gencode arch=compute_100a,code=sm_100a -cubin -Xptxas -v -fmad=false out.cu` shows registers / spills,
and `cuobjdump -sass` the instruction mix, before any GPU time is spent.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import descent_b200 as d

network, mini_batch, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
env = d.Environment(device=-1)
ex = d.Example(env, network, mini_batch, "adam", 0.0, 512, 512)
open(out, "w").write(ex.train_graph.kernel_source(tf32="--strict" not in sys.argv))
