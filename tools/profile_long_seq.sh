#!/bin/bash
# ncu --set full of the long-sequence kernels: attention forward at b32 x 1025 tokens, flash attention backward at the same shape
set -u
tag=${1:-r2e}
out=gpurun_out
mkdir -p $out
python tools/attn_bench.py > $out/plain_attn_$tag.log 2>&1 || { echo "plain attn_bench failed"; exit 1; }
python tools/train_profile.py 32 512 > $out/plain_train512_$tag.log 2>&1 || { echo "plain train_profile failed"; tail -3 $out/plain_train512_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:attention_v3 -s 46 -c 1 -o $out/attn512_$tag python tools/attn_bench.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_bwd_flash -s 2 -c 1 -o $out/bwdflash512_$tag python tools/train_profile.py 32 512 > /dev/null 2>&1
ls -la $out | grep $tag
