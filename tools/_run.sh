set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke1.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1_fp16.json 2> gpurun_out/r2_bench1_fp16.err
cp gpurun_out/layers_b56_fp16.json gpurun_out/r2_layers1_fp16.json
python bench.py --steps 3 --warmup 3 --dtype bf16 --lean > gpurun_out/r2_bench1_bf16.json 2> gpurun_out/r2_bench1_bf16.err
cp gpurun_out/layers_b56_bf16.json gpurun_out/r2_layers1_bf16.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench1_ref.json 2> gpurun_out/r2_bench1_ref.err
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -k regex:"igemm_tc|stem_tc" -o gpurun_out/r2_traffic_b56 -f python tools/one_forward.py 56 2 > gpurun_out/r2_ncu_traffic.log 2>&1
ls -la gpurun_out | tail -5
