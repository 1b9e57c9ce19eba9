# experiment: where do the extra CG iterations of the partitioned mode come from (in-process partitions on one GPU)
for nx in 1000 2000; do
echo "== nx $nx default";                 python tests/tools/gpu_partition_iterations.py $nx 2>&1 | tail -2
echo "== nx $nx replicate all below L0";  PARTITION_ONLY=1 ARAP_MG_REPLICATE_ROWS=100000000 python tests/tools/gpu_partition_iterations.py $nx 2>&1 | tail -1
echo "== nx $nx replicate none";          PARTITION_ONLY=1 ARAP_MG_REPLICATE_ROWS=0 python tests/tools/gpu_partition_iterations.py $nx 2>&1 | tail -1
echo "== nx $nx host setup";              ARAP_MG_DEVICE_SETUP=0 python tests/tools/gpu_partition_iterations.py $nx 2>&1 | tail -2
echo "== nx $nx rim keys";                ARAP_MG_AGG_KEY=rim python tests/tools/gpu_partition_iterations.py $nx 2>&1 | tail -2
done
