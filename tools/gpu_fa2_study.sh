#!/bin/bash
# v2 fused attention study on the GPU box: elimination variants (build/variants/lib_fa2_dbg*.so) and the role timeline
tag=${1:-t2}
out=gpurun_out
mkdir -p $out
: > $out/${tag}_times.jsonl
for so in build/variants/lib_fa2_dbg*.so; do
  HMVIT_LIB=$PWD/$so TIME_ATTN_IMPLS=fused timeout 120 python tools/time_attn.py >> $out/${tag}_times.jsonl 2>>$out/${tag}.err
done
cat $out/${tag}_times.jsonl
for k in 0 1; do
HMVIT_LIB=$PWD/build/variants/lib_fa2_ts.so TS_ABS=1 TS_T0=300000 TS_T1=340000 timeout 120 python tools/fa_timeline.py $k > $out/${tag}_timeline$k.txt 2>>$out/${tag}.err
done
tail -n 5 $out/${tag}.err
