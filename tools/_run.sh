set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_pytest9.log
python tools/layer_ab.py 56 4 "" "" > gpurun_out/r2_layer_ab5.txt 2>&1
