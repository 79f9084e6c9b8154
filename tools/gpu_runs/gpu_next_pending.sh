#!/bin/bash
# What was NOT yet run on a GPU when round 2 ended (profiles/r02_notes.md section 8), in the order to run it.  Each step is bounded.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_runs/gpu_next_pending.sh 2'      (2 GPUs: protocol + parity)
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_runs/gpu_next_pending.sh 8'      (8 GPUs: scaling figures incl. numRefs 6)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# 1. single-GPU suite (VTK test of the device-resident function is new)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/next_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/next_pytest.log
# 2. decomposed vs single GPU with the capped exchange grid and the exact Gershgorin bound: identical BiCGStab counts expected at every N
timeout 300 $TR --master-port 29731 tools/dist_check.py 3 3 7000 > gpurun_out/next_dist3d_n$N.json 2> gpurun_out/next_dist3d_n$N.err; echo "dist_check 3D rc=$?"; tail -c 600 gpurun_out/next_dist3d_n$N.json
timeout 300 $TR --master-port 29732 tools/dist_check.py 5 2 1500 > gpurun_out/next_dist2d_n$N.json 2> gpurun_out/next_dist2d_n$N.err; echo "dist_check 2D rc=$?"; tail -c 600 gpurun_out/next_dist2d_n$N.json
# 3. the bench line at N (parity.decomposed should now show equal BiCGStab counts; admm_refs4 at N = 8 was 348 ms with the loose bound)
timeout 600 $TR --master-port 29733 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/next_bench_n$N.json 2> gpurun_out/next_bench_n$N.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/next_bench_n$N.json
# 4. the largest refinement (the interface of level 6 needs more CTAs than fit: the case the capped grid was written for)
if [ "$N" -ge 8 ]; then
  timeout 300 $TR --master-port 29734 tools/scale_large.py solver 6 > gpurun_out/next_solver_n${N}_r6.json 2> gpurun_out/next_solver_n${N}_r6.err; echo "solver6 rc=$?"; cat gpurun_out/next_solver_n${N}_r6.json
  timeout 400 $TR --master-port 29735 tools/scale_large.py admm 6 > gpurun_out/next_admm_n${N}_r6.json 2> gpurun_out/next_admm_n${N}_r6.err; echo "admm6 rc=$?"; cat gpurun_out/next_admm_n${N}_r6.json
fi
