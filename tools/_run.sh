set -x
python tools/layer_ab.py 56 4 "" "fixcc=32" "fixcc=16" "pair=15" "" > gpurun_out/r2_layer_ab2.txt 2>&1
