#!/bin/bash
# multi-GPU session: parity of the partitioned boxes (log kept), the multi-GPU pytest cases, weak and strong scaling lines
#   usage: gpu_mgpu.sh N    (run under gpurun --gpus N)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
NSDG_MGPU_LOG=gpurun_out/r2_mgpu_parity_n$N.txt run 29511 tests/mgpu_parity.py > gpurun_out/r2_mgpu_parity_n$N.out 2>&1; tail -3 gpurun_out/r2_mgpu_parity_n$N.out
if [ "$N" = "2" ]; then
  (timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "multi_gpu or two_devices") > gpurun_out/r2_mgpu_pytest_n$N.log 2>&1; tail -3 gpurun_out/r2_mgpu_pytest_n$N.log
fi
run 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_weak_n$N.json 2> gpurun_out/r2_bench_weak_n$N.err; tail -2 gpurun_out/r2_bench_weak_n$N.err; head -c 300 gpurun_out/r2_bench_weak_n$N.json; echo
run 29513 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong > gpurun_out/r2_bench_strong_n$N.json 2> gpurun_out/r2_bench_strong_n$N.err; tail -2 gpurun_out/r2_bench_strong_n$N.err; head -c 300 gpurun_out/r2_bench_strong_n$N.json; echo
