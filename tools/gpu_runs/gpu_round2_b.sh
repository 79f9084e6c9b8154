#!/bin/bash
# 1-GPU pass: full GPU test-suite, full bench line, 2D SpMV and assembly A/B, ncu DRAM bytes of one V-cycle at numRefs 5
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02b_pytest.log
tail -4 gpurun_out/r02b_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02b_bench_n1.json; tail -3 gpurun_out/r02b_bench_n1.err
for lanes in 8 4; do
  ADMM_B200_SPMV2D_LANES=$lanes timeout 300 python tools/scale_large.py solver 7 2 > gpurun_out/r02b_2d_r7_lanes$lanes.json 2> gpurun_out/r02b_2d_lanes$lanes.err; echo "2d lanes=$lanes rc=$?"; cat gpurun_out/r02b_2d_r7_lanes$lanes.json
done
for asm in atomic rows; do
  ADMM_B200_ASSEMBLY=$asm timeout 600 python tools/scale_large.py solver 6 > gpurun_out/r02b_solver_r6_$asm.json 2> gpurun_out/r02b_solver_r6_$asm.err; echo "solver6 asm=$asm rc=$?"; cat gpurun_out/r02b_solver_r6_$asm.json
done
timeout 300 python tools/vcycle_once.py 5 2 > gpurun_out/r02b_vcycle_plain.log 2>&1 && \
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_vcycle_L5_dram.csv python tools/vcycle_once.py 5 2 > gpurun_out/r02b_vcycle_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r02b_vcycle_plain.log; wc -l gpurun_out/r02b_vcycle_L5_dram.csv
