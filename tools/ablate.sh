#!/bin/bash
# dev aid: true marginal cost of each op family inside the overlapped schedule = step time with the family's launches removed
# (XF_ABLATE, transfusion_b200/ops.py).  usage (GPU box): tools/ablate.sh [families...]
fams=${@:-"none layernorm_fwd layernorm_bwd colsum patchify_fold attn_delta rows_gather cast attn_fwd attn_bwd gemm_fwd gemm_dgrad gemm_wgrad"}
for f in $fams; do
  XF_ABLATE=$([ "$f" = none ] && echo "" || echo "$f") timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bf16-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
