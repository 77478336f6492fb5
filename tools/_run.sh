set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_fp16.json 2> gpurun_out/r02_bench_fp16.err
cp gpurun_out/layers_b56_fp16.json gpurun_out/r02_layers_b56_fp16.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --split sweep --steps 5 --warmup 2 > gpurun_out/r02_split_sweep_2gpu.json 2> gpurun_out/r02_split_sweep_2gpu.err
