j() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
for i in 1 2; do
echo impl0; python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | j
echo impl3; LDIT_ATTN_IMPL=3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | j
echo impl3 p1q2; LDIT_LIB_PATH=/root/repo/build_variants/lib_p1q2.so LDIT_ATTN_IMPL=3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | j
done
echo 512 impl0; python bench.py --workload base512 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | j
echo 512 impl3; LDIT_ATTN_IMPL=3 python bench.py --workload base512 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | j
echo 512 impl3 p1q2; LDIT_LIB_PATH=/root/repo/build_variants/lib_p1q2.so LDIT_ATTN_IMPL=3 python bench.py --workload base512 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | j
LDIT_ATTN_IMPL=3 python tools/step_profile.py base224 | grep -E "attention|eager"
LDIT_ATTN_IMPL=3 python tools/step_profile.py base512 | grep -E "attention|eager"
