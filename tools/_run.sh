set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2_pytest3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke3.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench3_fp16.json 2> gpurun_out/r2_bench3_fp16.err
cp gpurun_out/layers_b56_fp16.json gpurun_out/r2_layers3_fp16.json
python tools/call_patterns.py gpurun_out/r02_call_patterns.json > gpurun_out/r2_call_patterns.log 2>&1
python tools/batch_sweep.py gpurun_out/r02_batch_sweep.json > gpurun_out/r2_batch_sweep.log 2>&1
