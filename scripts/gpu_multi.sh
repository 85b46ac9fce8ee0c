#!/bin/bash
# multi-GPU validation on an N-GPU box: bash scripts/gpu_multi.sh N [reps]
N=${1:-2}; REPS=${2:-1}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    timeout 120 $TR --nproc-per-node $n --master-port $((29600+n)) scripts/h2d_probe_ranks.py 2>gpurun_out/h2d_${n}.err | grep h2d_probe | tee -a gpurun_out/h2d_probe_ranks_${N}gpu.txt
  fi
done
for r in $(seq 1 $REPS); do
  timeout 400 $TR --nproc-per-node $N --master-port $((29700+r)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu_run${r}.json 2> gpurun_out/bench_${N}gpu_run${r}.err
  echo "run $r rc=$?"; tail -c 1500 gpurun_out/bench_${N}gpu_run${r}.json; tail -3 gpurun_out/bench_${N}gpu_run${r}.err
done
timeout 400 $TR --nproc-per-node $N --master-port 29800 bench.py --gpus $N --steps 10 --warmup 3 --config 5 > gpurun_out/bench_${N}gpu_cfg5.json 2> gpurun_out/bench_${N}gpu_cfg5.err
echo "cfg5 rc=$?"; tail -c 1200 gpurun_out/bench_${N}gpu_cfg5.json; tail -3 gpurun_out/bench_${N}gpu_cfg5.err
