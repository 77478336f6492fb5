set -x
python tools/layer_ab.py 56 4 "" "rs=0" "ctas=3" "" > gpurun_out/r2_layer_ab3.txt 2>&1
