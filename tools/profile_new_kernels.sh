#!/bin/bash
# `ncu --set full` captures of the kernels added late in round 2: attention forward with a relative-position table,
# attention backward (tcgen05), the wgrad / dgrad GEMM modes, LayerNorm backward.  usage: bash tools/profile_new_kernels.sh <tag>
set -u
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
python tools/attn_bench.py > $out/plain_attn_$tag.log 2>&1 || { echo "plain attn_bench failed"; tail -5 $out/plain_attn_$tag.log; exit 1; }
python tools/train_bench.py 1 > $out/plain_train_$tag.log 2>&1 || { echo "plain train_bench failed"; tail -5 $out/plain_train_$tag.log; exit 1; }
# attn_bench launches per shape: 23 without a table, then 23 with one -> launch 23 is the first biased launch at base224
ncu --set full --clock-control none --import-source on -k regex:attention_v3 -s 23 -c 1 -o $out/attnbias_$tag python tools/attn_bench.py > /dev/null 2>&1
# one layer forward + backward: GEMM launches 0-3 forward (qkv, proj, fc1, fc2), 4-11 backward (wgrad / dgrad alternating)
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 4 -c 8 -o $out/bwdgemm_$tag python tools/train_bench.py 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attention_bwd|layernorm_bwd|scale_residual_bwd|gelu_bwd" -c 6 -o $out/bwdrow_$tag python tools/train_bench.py 1 > /dev/null 2>&1
ls -la $out | grep $tag
