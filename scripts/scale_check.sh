#!/bin/bash
# usage (under gpurun --gpus N): scripts/scale_check.sh <tag> <N> [extra bench args]   -> gpurun_out/<tag>_bench_n{1,N}.json + efficiency
TAG=$1; N=$2; shift 2
python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline "$@" 2>gpurun_out/${TAG}_n1.err | grep '^{' > gpurun_out/${TAG}_bench_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-extras "$@" 2>gpurun_out/${TAG}_n${N}.err | grep '^{' > gpurun_out/${TAG}_bench_n${N}.json
tail -3 gpurun_out/${TAG}_n${N}.err
python - <<PY
import json
a=json.load(open("gpurun_out/${TAG}_bench_n1.json")); b=json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
print("${TAG}: N=1 %.4g samples/s %.4f ms | N=${N} %.4g samples/s %.4f ms | efficiency %.4f | exposed %.1f us" % (a["value"], a["ms_per_step"], b["value"], b["ms_per_step"], b["value"]/${N}/a["value"], (b["ms_per_step"]-a["ms_per_step"])*1e3))
PY
