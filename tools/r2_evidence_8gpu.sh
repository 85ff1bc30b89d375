#!/bin/bash
# Round-2 multi-GPU evidence on ONE 8-GPU box (gpurun --gpus 8): data-parallel checks at world 8, the headline at its
# real length (1500 steps) with the secondary block, BASELINE config 5 at T = 1500.  Every command has its own timeout.
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 tests/multi/dp_worker.py > $O/r2_dp_worker_${N}gpu.log 2>&1; echo "dp_worker rc=$?"; grep "DP-WORKER\|Error\|assert" $O/r2_dp_worker_${N}gpu.log | head -5
timeout 420 $TR --master-port 29602 bench.py --gpus $N --steps 1500 --warmup 5 > $O/r2_bench_full1500_${N}gpu.log 2>&1; echo "bench rc=$?"; tail -1 $O/r2_bench_full1500_${N}gpu.log | cut -c1-400
timeout 300 $TR --master-port 29603 tools/nll_sharded.py 4096 1500 512 > $O/r2_nll_sharded_${N}gpu.log 2>&1; echo "nll rc=$?"; grep NLL-SHARDED $O/r2_nll_sharded_${N}gpu.log
