set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke2.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2_fp16.json 2> gpurun_out/r2_bench2_fp16.err
cp gpurun_out/layers_b56_fp16.json gpurun_out/r2_layers2_fp16.json
bash tools/ab.sh "--opt pdl=1" "--opt ng=4" 1 > gpurun_out/r2_ab_ng4.txt 2>&1
bash tools/ab.sh "--opt keep_sum=0 --opt stem_lo=0" "--batch 84" 1 > gpurun_out/r2_ab_prec_b84.txt 2>&1
AB_BATCH=56 bash tools/ab.sh "--batch 56" "--batch 56 --opt ctas=1" 1 > gpurun_out/r2_ab_ctas1.txt 2>&1
