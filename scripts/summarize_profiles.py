"""Turn the outputs of scripts/profile_round.sh (gpurun_out/<tag>_<workload>_*) into the tracked evidence under profiles/.

usage: python scripts/summarize_profiles.py <tag> <workload> <mini_batch>

Writes
  profiles/<tag>_<workload>_launches.md   per-kernel share of the step: ncu launch list vs CUDA-event timings
  profiles/<tag>_<workload>_ncu_full.md   one row per kernel from the `ncu --set full` capture (DRAM bytes, pipe use, stalls)
and merges the measured DRAM bytes per launch into profiles/dram_traffic.json (read by bench.py for `roofline.traffic`).
Runs here on the CPU box: `ncu -i` only reads the report.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, workload, mini_batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
base = os.path.join(ROOT, "gpurun_out", "%s_%s" % (tag, workload))
events = json.load(open(base + "_events.json"))["launches"]
bench = json.load(open(base + "_bench.json"))
label_of = {}
event_ms = {}
for l in events:
    if l["entry"] == "dsc_gemm_tf32":
        continue
    label_of[l["entry"]] = l["label"]
    event_ms[l["entry"]] = event_ms.get(l["entry"], 0.0) + l["ms"]
dense_ms = sum(l["ms"] for l in events if l["entry"] == "dsc_gemm_tf32")
event_total = sum(l["ms"] for l in events)

# ---- launch list ---------------------------------------------------------------------------------
rows = [r for r in csv.reader(l for l in open(base + "_launches.csv") if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ncu_ns, ncu_count = {}, {}
for r in rows[1:]:
    name = "dsc_gemm_tf32" if "gemm_tf32" in r[ki] else r[ki]
    ncu_ns[name] = ncu_ns.get(name, 0.0) + float(r[vi])
    ncu_count[name] = ncu_count.get(name, 0) + 1
ncu_total = sum(ncu_ns.values())
counts = sorted(c for n, c in ncu_count.items() if n != "dsc_gemm_tf32")
steps = counts[len(counts) // 2] if counts else 1  # launches of a kernel that runs once per step
out = ["# %s - %s, mini-batch %d, 1x B200: per-kernel launch list" % (tag, workload, mini_batch), "",
       "Command (under gpurun): `scripts/profile_round.sh %s %s` -> `ncu --metrics gpu__time_duration.sum --clock-control none` over %d replays of the"
       " training step's CUDA graph (ncu times are cold-cache and serialised: compare SHARES)." % (tag, workload, steps),
       "`events share` is the same kernel's share of an un-profiled pass timed with CUDA events inside `bench.py --profile-json`.", "",
       "| kernel | label | launches | ncu us / step | ncu share | events us / step | events share |", "|---|---|---|---|---|---|---|"]
for name, ns in sorted(ncu_ns.items(), key=lambda kv: -kv[1]):
    ev = dense_ms if name == "dsc_gemm_tf32" else event_ms.get(name, 0.0)
    if ns / ncu_total < 0.002:
        continue
    out.append("| %s | %s | %d | %.1f | %.1f%% | %.1f | %.1f%% |" % (name, label_of.get(name, "TMA-fed dense tcgen05 TF32 GEMM (precompiled)" if "gemm_tf32" in name else "runtime helper (seed store / fill / one-shot all-reduce)"), ncu_count[name],
                                                                 ns / 1e3 / steps, 100 * ns / ncu_total, ev * 1e3, 100 * ev / event_total))
out += ["", "Totals: ncu %.2f ms/step (serialised, cold); CUDA-event eager pass %.2f ms/step; CUDA-graph replay measured by bench.py: %.2f ms/step "
        "(%.0f samples/s)." % (ncu_total / 1e6 / steps, event_total, bench["ms_per_step"], bench["value"])]
open(os.path.join(ROOT, "profiles", "%s_%s_launches.md" % (tag, workload)), "w").write("\n".join(out) + "\n")

# ---- full capture --------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", base + "_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name):
    try:
        return float(r[col[name]])
    except (KeyError, ValueError):
        return float("nan")


seen = set()
lines = ["# %s - %s, mini-batch %d: `ncu --set full --clock-control none --import-source on` per kernel" % (tag, workload, mini_batch), "",
         "One launch of every kernel of the training step (first occurrence).  `DRAM MB` = dram__bytes_read.sum / dram__bytes_write.sum;"
         " `alg MB` = the algorithmic bytes bench.py uses for the roofline (SURVEY.md section 8d); `tensor %` = sm__pipe_tensor_cycles_active;"
         " `tc-smem %` = l1tex__data_pipe_tc_wavefronts_mem_shared (tensor-core operand reads from shared memory);"
         " stalls are warps-per-issue ratios above 1.", "",
         "| kernel | label | us | DRAM MB r/w | alg MB | dram % | sm % | tensor % | tc-smem % | issue % | warps % | regs | top stalls |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
alg = {l["entry"]: l["bytes"] for l in events}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if "gemm_tf32" in name:
        name = "dsc_gemm_tf32"
    us = f(r, "gpu__time_duration.sum")
    if name in seen and name != "dsc_gemm_tf32":
        continue
    seen.add(name)
    if us < 15:
        continue
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            key = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
            if v > 1.0 and key != "selected":
                stalls.append((v, key))
    stalls.sort(reverse=True)
    label = label_of.get(name, "dense TF32 GEMM")
    lines.append("| %s | %s | %.1f | %.0f / %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %d | %s |" % (
        name, label, us, rd, wr, alg.get(name, float("nan")) / 1e6, f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        f(r, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), int(f(r, "launch__registers_per_thread")),
        ", ".join("%s %.1f" % (k, v) for v, k in stalls[:3])))
    if name != "dsc_gemm_tf32":
        traffic["%s/%d/%s" % (workload, mini_batch, label)] = int((rd + wr) * 1e6)
open(os.path.join(ROOT, "profiles", "%s_%s_ncu_full.md" % (tag, workload)), "w").write("\n".join(lines) + "\n")

path = os.path.join(ROOT, "profiles", "dram_traffic.json")
merged = json.load(open(path)) if os.path.exists(path) else {}
merged["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full captures (profiles/*_ncu_full.md); "
                      "key = workload/mini_batch/kernel label")
merged.update(traffic)
json.dump(merged, open(path, "w"), indent=1, sort_keys=True)
print("wrote profiles/%s_%s_{launches,ncu_full}.md and %d traffic entries" % (tag, workload, len(traffic)))
