#!/bin/bash
# One-box evidence run (gpurun): full GPU test suite, bench (ours + reference arm), isolated kernels, torch profile, ncu launch list and
# a --set full capture of the attention / KL / router kernels.  Outputs under gpurun_out/; the summaries are copied into profiles/ by hand.
TAG=${1:-r2_final}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gputest_$TAG.log 2>&1; tail -4 gpurun_out/gputest_$TAG.log
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; tail -2 gpurun_out/bench_${TAG}_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
REPS=20 timeout 300 python profiles/microbench.py > gpurun_out/microbench_$TAG.txt 2>&1
timeout 300 python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e --torch-profile gpurun_out/torch_profile_$TAG.txt > /dev/null 2> gpurun_out/torch_profile_$TAG.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd|attn_bwd|kl_stream|moe_gate|moe_seat' -c 16 -o gpurun_out/kernels_$TAG -f python profiles/microbench.py > gpurun_out/ncu_kernels_$TAG.log 2>&1; tail -2 gpurun_out/ncu_kernels_$TAG.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "kl", d["roofline"]["frac"], "launches", d["gpu_launches"], "clocks", d["clocks"])
for k,v in (d.get("secondary") or {}).items(): print(k, v.get("value"), v.get("unit"), v.get("ms_per_step"))
r=json.load(open("gpurun_out/bench_${TAG}_ref.json")); print("ref", r["value"], r["ms_per_step"])
PY
