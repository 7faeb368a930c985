#!/bin/bash
# compute-sanitizer over one launch of every hot kernel (tools/profile_kernels.py at a reduced batch / sequence so the
# instrumented run finishes in minutes).  Summaries go to gpurun_out/; copy them to profiles/ to keep them.
#   usage (GPU box): tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]
mkdir -p gpurun_out
export XF_B=${XF_B:-1} XF_S=${XF_S:-832}
for tool in "${@:-memcheck racecheck}"; do
  for t in $tool; do
    echo "== compute-sanitizer --tool $t (XF_B=$XF_B XF_S=$XF_S)"
    timeout 900 compute-sanitizer --tool "$t" --print-limit 20 python tools/profile_kernels.py > "gpurun_out/sanitizer_$t.log" 2>&1
    echo "rc=$?" >> "gpurun_out/sanitizer_$t.log"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|profile_kernels ok|rc=|Error|hazard" "gpurun_out/sanitizer_$t.log" | sort | uniq -c | head -20
  done
done
