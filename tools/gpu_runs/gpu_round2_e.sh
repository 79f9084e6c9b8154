#!/bin/bash
# 8-GPU pass (bounded): full bench line at N=8, largest refinement (numRefs 6): solver figures and one full ADMM iteration
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29721 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02e_bench_n8.json 2> gpurun_out/r02e_bench_n8.err; echo "bench8 rc=$?"
tail -c 3800 gpurun_out/r02e_bench_n8.json; grep -E "Error|error|assert" gpurun_out/r02e_bench_n8.err | head -5
timeout 240 $TR --master-port 29722 tools/scale_large.py solver 6 > gpurun_out/r02e_solver_n8_r6.json 2> gpurun_out/r02e_solver_n8_r6.err; echo "solver6 rc=$?"; cat gpurun_out/r02e_solver_n8_r6.json
timeout 300 $TR --master-port 29723 tools/scale_large.py admm 6 > gpurun_out/r02e_admm_n8_r6.json 2> gpurun_out/r02e_admm_n8_r6.err; echo "admm6 rc=$?"; cat gpurun_out/r02e_admm_n8_r6.json
