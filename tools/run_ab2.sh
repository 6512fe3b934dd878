j() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'])"; }
export LDIT_LIB_PATH=/root/repo/build_variants/lib_exp.so
for WL in base224 base512; do
for i in 1 2 3; do
  echo -n "$WL attn v3 : "; LDIT_ATTN_IMPL=0 python bench.py --workload $WL --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | j
  echo -n "$WL attn old: "; LDIT_ATTN_IMPL=4 python bench.py --workload $WL --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | j
done; done
LDIT_ATTN_IMPL=0 python tools/step_profile.py base224 | head -9
LDIT_ATTN_IMPL=4 python tools/step_profile.py base224 | head -9
