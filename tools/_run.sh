set -x
python -m pytest tests -m gpu -q -x -k "fused_transposed or planner_variants or small_and_odd or golden" 2>&1 | tail -6 > gpurun_out/r2_pytest7.log
python tools/layer_ab.py 56 4 "fixcompact=0" "" "fixcompact=0" "" > gpurun_out/r2_layer_ab4.txt 2>&1
