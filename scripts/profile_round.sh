#!/bin/bash
# Evidence pass for one round (run under gpurun, one GPU):  scripts/profile_round.sh <tag> [workload] [kernel-regex]
#   1. launch list of two training steps with per-launch device time (shares only: cold cache, serialised)
#   2. one `ncu --set full` capture of the kernels matching the regex (default: the JIT kernels k<N> and the dense GEMM)
#   3. an unprofiled bench run with per-kernel CUDA-event timings for the share comparison
# Outputs land in gpurun_out/; scripts/summarize_profiles.py turns them into profiles/*.md here on the CPU box.
set -u
TAG=${1:-r1}
WORKLOAD=${2:-conv-net}
REGEX=${3:-'^(k[0-9]+|.*gemm_tf32.*)$'}
mkdir -p gpurun_out
python bench.py --workload "$WORKLOAD" --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/${TAG}_${WORKLOAD}_events.json > gpurun_out/${TAG}_${WORKLOAD}_bench.json 2> gpurun_out/${TAG}_${WORKLOAD}_bench.err
KPS=$(python -c "import json;print(json.load(open('gpurun_out/${TAG}_${WORKLOAD}_bench.json'))['kernels_per_step'])")
# bench.py replays warm-up + timed steps through one CUDA graph; skip the eager JIT/first step, keep two steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((KPS * 3)) -c $((KPS * 2)) --csv \
    --log-file gpurun_out/${TAG}_${WORKLOAD}_launches.csv python bench.py --workload "$WORKLOAD" --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $((KPS * 3)) -c $KPS \
    -o gpurun_out/${TAG}_${WORKLOAD}_full -f python bench.py --workload "$WORKLOAD" --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${WORKLOAD}_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
