#!/bin/bash
# 2-GPU validation pass (bounded): multi-GPU parity tests, full bench line at N=2
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -k "multi_gpu or undivided" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02d_pytest.log
tail -4 gpurun_out/r02d_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err; echo "bench2 rc=$?"
tail -c 3500 gpurun_out/r02d_bench_n2.json; grep -E "Error|error|assert" gpurun_out/r02d_bench_n2.err | head -5
