# same-box A/B of the current library against a variant ($1, under build_variants/), base224 and base512
j() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
V=$1
python -m pytest tests/test_kernels_gpu.py -x -q -k attention 2>&1 | tail -2
for i in 1 2 3; do
  echo -n "224 current : "; python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | j
  echo -n "224 $V: "; LDIT_LIB_PATH=/root/repo/build_variants/$V python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | j
done
for i in 1 2; do
  echo -n "512 current : "; python bench.py --workload base512 --steps 15 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | j
  echo -n "512 $V: "; LDIT_LIB_PATH=/root/repo/build_variants/$V python bench.py --workload base512 --steps 15 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | j
done
python tools/step_profile.py base224 | grep -E "attention|eager"
python tools/step_profile.py base512 | grep -E "attention|eager"
