#!/bin/bash
# first GPU contact of a new kernel: a few parity checks in isolated subprocesses (a trapped kernel poisons only its
# process), then a short bench.  Usage: gpurun --timeout 900 -- 'bash tools/gpu_r2_first.sh <tag> [check ...]'
tag=${1:-r2a}; shift
out=gpurun_out
mkdir -p $out
checks=${@:-attn_split_vs_single attention_golden fusion_small fusion_golden ragged_batch fusion_config2_scene attn_bwd}
BRINGUP_TIMEOUT=150 timeout 800 python tools/bringup.py $checks > $out/${tag}_bringup.log 2>&1
cat $out/${tag}_bringup.log | cut -c1-600
timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 400 $out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
    print(json.dumps(d["kernels"]))
except Exception as e:
    print("bench parse failed", e)
PY
