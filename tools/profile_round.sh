#!/bin/bash
# One gpurun call: launch list of a bench step + `ncu --set full` captures of each kernel family.
# usage (on the GPU box, from the repo root): bash tools/profile_round.sh <tag>
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$cmd > $out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/launches_$tag.csv $cmd > $out/ncu_launches_$tag.log 2>&1
# GEMMs of the second forward: QKV, proj (bf16 branch), fc1, fc2 (48 GEMM launches per forward + the patch embed)
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 49 -c 5 -o $out/gemm_$tag $cmd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_v3 -s 12 -c 1 -o $out/attn_$tag $cmd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:layernorm -s 24 -c 2 -o $out/ln_$tag $cmd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"im2col|taps" -s 5 -c 5 -o $out/rowwise_$tag $cmd > /dev/null 2>&1
# TMA-fed patch embedding (fp16 pages), FPN head (row f1), fused input transform (row f3)
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 8 -c 1 -o $out/patchtma_$tag python tools/patch_bench.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_fpn_$tag.csv python tools/fpn_one.py > $out/ncu_launches_fpn_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o $out/conv_$tag python tools/conv_one.py > /dev/null 2>&1
ls -la $out | tail -12
