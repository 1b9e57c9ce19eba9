#!/bin/bash
# scripts/gpu_profile.sh -- full-size bench + ncu launch list + full captures of the hot kernels.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
if [ -z "$NO_NCU" ]; then
# launch list (cold-cache, serialised): shares only. CUDA graphs are disabled by the per-launch profile? no: ncu sees graph kernel nodes too.
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-2800} -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"local_step|rhs_residual" \
    -s 2 -c 4 -o gpurun_out/prof_local -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_local.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cg_spmv|cg_update_mg|mg_fine|mg_restrict|mg_csr" \
    -s 40 -c 16 -o gpurun_out/prof_hot -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
