#!/bin/bash
# One gpurun call that refreshes every piece of measured evidence of a build (run from the repo root ON the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round_check.sh r2'
# Writes gpurun_out/<tag>_{tests.log,smoke.log,bench.json,launches.csv,ncu_full.log,prof.ncu-rep}; summarise them here with
#   python tools/ncu_launches.py gpurun_out/<tag>_launches.csv "<title>" "<command>" > profiles/<tag>_launches.md
#   python tools/ncu_summary.py  gpurun_out/<tag>_prof.ncu-rep  "<title>" "<command>" > profiles/<tag>_ncu_full.md
#   python tools/ncu_traffic.py  gpurun_out/<tag>_prof.ncu-rep      # rewrites profiles/r2_traffic.json (bench.py reads it)
# Numbers printed under ncu are never bench values; the bench line comes from the un-profiled run.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
timeout 60 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-train --no-stress --no-cpu-baseline > $out/${tag}_ncu_b.log 2>&1
# full-set capture of the forward's kernels after the first forward's launches (the command round 1's captures used)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fused_attn|tap_records|qkv_kernel|chain_kernel" \
    -s 13 -c 14 -o $out/${tag}_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-stress > $out/${tag}_ncu_full.log 2>&1
tail -n 3 $out/${tag}_tests.log $out/${tag}_smoke.log
head -c 600 $out/${tag}_bench.json; echo
