# same-box A/B of the three p5 gather modes at N GPUs ($1)
N=${1:-2}
j() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], '|', d['config']['gather'][:60])"; }
for m in peer nccl nccl-side peer nccl; do
  echo -n "$m: "; LDIT_BENCH_GATHER=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-extras --no-cpu-baseline 2>gpurun_out/gather_$m.err | j
  grep -h "unavailable" gpurun_out/gather_$m.err | head -2
done
