#!/bin/bash
# 2-GPU pass: multi-GPU parity tests, full bench line at N=2, agglomeration-threshold A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "multi_gpu or undivided or trace_files" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02c_pytest.log
tail -6 gpurun_out/r02c_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err; echo "bench2 rc=$?"
tail -c 2500 gpurun_out/r02c_bench_n2.json; tail -4 gpurun_out/r02c_bench_n2.err
for refs in 5 4; do for g in 400000 50000; do
  ADMM_B200_GATHER_DOFS=$g timeout 300 $TR --master-port 29712 tools/scale_large.py solver $refs > gpurun_out/r02c_solver_n2_r${refs}_g$g.json 2> gpurun_out/r02c_solver_n2_r${refs}_g$g.err; echo "solver refs=$refs gather=$g rc=$?"; cat gpurun_out/r02c_solver_n2_r${refs}_g$g.json
done; done
