# bash tools/run_multi_gpu.sh N  (on a box with N GPUs): frame-sharded single sweep, weak-scaling bench and the CPU reference arm under torchrun
set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --split sweep --steps 5 --warmup 2 > gpurun_out/r02_split_sweep_${N}gpu.json 2> gpurun_out/r02_split_sweep_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_${N}gpu.json 2> gpurun_out/r02_bench_ref_${N}gpu.err
