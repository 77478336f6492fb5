set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_pytest8.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke8.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_fp16.json 2> gpurun_out/r02_bench_fp16.err
cp gpurun_out/layers_b56_fp16.json gpurun_out/r02_layers_b56_fp16.json
python bench.py --steps 5 --warmup 3 --dtype bf16 --lean > gpurun_out/r02_bench_bf16.json 2> gpurun_out/r02_bench_bf16.err
cp gpurun_out/layers_b56_bf16.json gpurun_out/r02_layers_b56_bf16.json
