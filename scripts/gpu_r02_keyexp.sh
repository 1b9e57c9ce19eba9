# experiment: wavefront (lexicographically first) aggregation on quad-like meshes
export COMPARE_DEVICE_ONLY=1 ARAP_MG_TIMING=1
python tests/tools/gpu_setup_compare.py grid:1000 grid:2000 grid:4000 2>&1 | grep -E "^device|device setup\] level|strong"
unset ARAP_MG_TIMING
for nx in 2000; do python tests/tools/gpu_partition_iterations.py $nx 2>&1 | tail -2; PARTITION_ONLY=1 python tests/tools/gpu_partition_iterations.py $nx 4 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "device_built or partitioned" 2>&1 | tail -3
