# experiment: what costs the partitioned hierarchy its quality -- numbering or the block constraint
T="python tests/tools/gpu_setup_compare.py grid:2000"
export COMPARE_DEVICE_ONLY=1
echo "== single default";             $T 2>&1 | grep -E "^device"
echo "== single, no renumbering";     ARAP_REORDER=0 $T 2>&1 | grep -E "^device"
echo "== single, no renumbering, index runs"; ARAP_REORDER=0 ARAP_MG_SWEEP_CELLS=0 $T 2>&1 | grep -E "^device"
echo "== single, 2 fake blocks";      ARAP_MG_FAKE_BLOCKS=2 $T 2>&1 | grep -E "^device"
echo "== single, 2 fake blocks, no renumbering"; ARAP_REORDER=0 ARAP_MG_FAKE_BLOCKS=2 $T 2>&1 | grep -E "^device"
echo "== single, 2 fake blocks, rim"; ARAP_MG_AGG_KEY=rim ARAP_MG_FAKE_BLOCKS=2 $T 2>&1 | grep -E "^device"
echo "== single, 2 fake blocks, host"; COMPARE_DEVICE_ONLY= ARAP_MG_FAKE_BLOCKS=2 $T 2>&1 | grep -E "^host"
