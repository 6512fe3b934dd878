timeout 300 python -m pytest tests/test_kernels_gpu.py -k "add_layernorm or gemm_bias_scale" -x -q 2>&1 | tail -3
LDIT_DEFER_RESID=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_fpn_gpu.py -x -q 2>&1 | tail -3
j() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
for WL in base224 base512 large224; do for i in 1 2 3; do
  echo -n "$WL resid-epilogue : "; LDIT_DEFER_RESID=0 python bench.py --workload $WL --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | j
  echo -n "$WL deferred add+LN: "; LDIT_DEFER_RESID=1 python bench.py --workload $WL --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | j
done; done
LDIT_DEFER_RESID=1 python tools/step_profile.py base224 | head -12
