#!/bin/bash
# round 2, multi-GPU pass: strong-scaling bench line at N GPUs (one stream sharded by contiguous block range) and the
# N-way host<->device copy ceiling of the box. usage: gpurun --gpus N -- bash scripts/gpu_r2_multi.sh N
N=${1:-8}
mkdir -p gpurun_out
export HSR_BENCH_TRACE=1
nvidia-smi topo -m > gpurun_out/r2_topo_n${N}.txt 2>&1
lscpu | grep -i -E "numa|model name|^cpu\(s\)|socket" >> gpurun_out/r2_topo_n${N}.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --headline-only --weak > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench_n${N}.json; grep -v "^\[bench\|^W\|^\*" gpurun_out/r2_bench_n${N}.err | tail -5
for k in 1 2 4 8; do
  if [ $k -le $N ]; then
    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2953$k \
        scripts/pcie_probe_nway.py >> gpurun_out/r2_pcie_probe_nway.jsonl 2>> gpurun_out/r2_pcie_probe_nway.err
  fi
done
cat gpurun_out/r2_pcie_probe_nway.jsonl | cut -c1-600
tail -3 gpurun_out/r2_pcie_probe_nway.err
