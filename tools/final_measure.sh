#!/bin/bash
# Round-end measurement batch on ONE GPU (run under gpurun): tests, the bench lines of every BASELINE config that fits one
# GPU, the reference arms, the config-5 sweep, the ncu launch list and the ncu --set full capture.  Outputs: gpurun_out/final_*.
mkdir -p gpurun_out
T="timeout 600"
python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final_tests.log; tail -3 gpurun_out/final_tests.log
b() { tag=$1; shift; $T python bench.py "$@" > gpurun_out/final_bench_$tag.json 2> gpurun_out/final_bench_$tag.err; echo "$tag rc=$? $(python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/final_bench_$tag.json").read().strip().splitlines()[-1])
    print(round(d.get("value",0),1), round(d.get("ms_per_step",0),2), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("whole_path_frac"))
except Exception as e:
    print("parse error", e)
PY
)"; }
b n1 --gpus 1 --steps 20 --warmup 5
b ego4dv1 --workload ego4dv1 --steps 10 --warmup 3 --no-cpu-baseline
b infer_b74 --mode infer --steps 10 --warmup 3 --no-cpu-baseline
b accum2 --accumulate 2 --steps 10 --warmup 4 --no-cpu-baseline --no-bf16-e2e
b laterals --with-fpn laterals --steps 10 --warmup 3 --no-cpu-baseline --no-bf16-e2e
XF_ATTN_BWD_WS=0 b 3pass --steps 10 --warmup 3 --no-cpu-baseline --no-bf16-e2e
b reference --impl reference --gpus 1 --steps 20 --warmup 5
b reference_gpu --impl reference-gpu --steps 5 --warmup 3
$T python bench.py --sweep --steps 4 --warmup 2 > gpurun_out/final_sweep.jsonl 2> gpurun_out/final_sweep.err; echo "sweep rc=$? lines=$(wc -l < gpurun_out/final_sweep.jsonl)"
# ncu: launch list of one bench step, then --set full of one launch of each hot kernel (tools/profile_kernels.py, 2nd iteration)
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-bf16-e2e > /dev/null 2>&1; echo "launch list rc=$?"
$T ncu --set full --clock-control none --import-source on -k regex:"tcgen05|layernorm|attn_delta" --launch-skip 13 -c 13 -f -o gpurun_out/final_full python tools/profile_kernels.py > gpurun_out/final_full.log 2>&1; echo "ncu full rc=$? $(ls -la gpurun_out/final_full.ncu-rep 2>/dev/null | awk '{print $5}')"
