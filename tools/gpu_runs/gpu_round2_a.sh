#!/bin/bash
# first GPU pass of round 2 (2-GPU box): GPU test-suite incl. the 2-rank parity checks, 1-GPU and 2-GPU bench smoke
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --roofline-refs 4 --no-cpu-baseline > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench1 rc=$?"
tail -c 1500 gpurun_out/r02a_bench_n1.json
for args in "2 3 100" "2 3 7000" "3 2 100"; do
  set -- $args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/dist_check.py $1 $2 $3 > gpurun_out/r02a_dist_$1_$2_$3.log 2>&1; echo "dist_check $args rc=$?"
  grep -E "rel L2|its|DIST CHECK|Error|error" gpurun_out/r02a_dist_$1_$2_$3.log | tail -8
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 5 --warmup 3 --roofline-refs 5 --no-cpu-baseline > gpurun_out/r02a_bench_n2.json 2> gpurun_out/r02a_bench_n2.err; echo "bench2 rc=$?"
tail -c 1800 gpurun_out/r02a_bench_n2.json; tail -5 gpurun_out/r02a_bench_n2.err
