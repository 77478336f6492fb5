set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_pytest5.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke5.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_fp16.json 2> gpurun_out/r02_bench_fp16.err
cp gpurun_out/layers_b56_fp16.json gpurun_out/r02_layers_b56_fp16.json
python bench.py --steps 5 --warmup 3 --dtype bf16 --lean > gpurun_out/r02_bench_bf16.json 2> gpurun_out/r02_bench_bf16.err
cp gpurun_out/layers_b56_bf16.json gpurun_out/r02_layers_b56_bf16.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 1 --warmup 3 --lean > gpurun_out/r02_ncu_launches_bench.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"igemm_tc|stem_tc" -o gpurun_out/r02_traffic_b56 -f python tools/one_forward.py 56 2 > gpurun_out/r02_ncu_traffic.log 2>&1
