#!/bin/bash
# attention study on the GPU box: elimination variants (build/variants/lib_fa_dbg*.so), role timelines (lib_fa_ts*.so)
# and, with a second argument, an ncu --set full capture of the fused kernel
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
: > $out/${tag}_attn_times.jsonl
timeout 120 python tools/time_attn.py >> $out/${tag}_attn_times.jsonl 2>$out/${tag}_attn.err
for so in build/variants/lib_fa_dbg*.so; do
  HMVIT_LIB=$PWD/$so TIME_ATTN_IMPLS=fused timeout 120 python tools/time_attn.py >> $out/${tag}_attn_times.jsonl 2>>$out/${tag}_attn.err
done
cat $out/${tag}_attn_times.jsonl
for so in build/variants/lib_fa_ts*.so; do
  n=$(basename $so .so)
  HMVIT_LIB=$PWD/$so timeout 120 python tools/fa_timeline.py 0 > $out/${tag}_${n}_timeline.txt 2>>$out/${tag}_attn.err
done
if [ -n "$2" ]; then
TIME_ATTN_IMPLS=fused timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fused_attn|tap_records" -s 82 -c 3 \
  -o $out/${tag}_fa -f python tools/time_attn.py > $out/${tag}_ncu.log 2>&1
tail -n 3 $out/${tag}_ncu.log
fi
tail -n 5 $out/${tag}_attn.err
