# same-box A/B: round-1 tree vs current library vs a variant library ($1), base224 (and base512 with $2=512)
j() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'])"; }
V=${1:-}
WL=${2:-base224}
for i in 1 2 3; do
  echo -n "r1      : "; (cd build_variants/r1 && python bench.py --workload $WL --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | j)
  echo -n "current : "; python bench.py --workload $WL --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | j
  if [ -n "$V" ]; then echo -n "$V: "; LDIT_LIB_PATH=/root/repo/build_variants/$V python bench.py --workload $WL --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | j; fi
done
