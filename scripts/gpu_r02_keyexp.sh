# experiment: device aggregation keys -- CG iterations, setup time, rounds per level
for k in ${CHUNKS:-6 8 10 12}; do echo "== chunk shift $k"
  COMPARE_DEVICE_ONLY=1 ARAP_MG_AGG_CHUNK=$k ARAP_MG_TIMING=1 python tests/tools/gpu_setup_compare.py ${MESHES:-grid:2000 ico:316} 2>&1 | grep -E "^device|^host|device setup\] level 0" 
done
