# Evidence run on the GPU box (gpurun -- bash tools/gpu_evidence.sh [prefix]): full GPU suite with printed
# measurements, bench lines, ncu launch lists and --set full captures; everything lands in gpurun_out/<prefix>_*.
P=${1:-r2b}
set -x
mkdir -p gpurun_out
# 1. whole GPU suite with the printed measurements (drift curves, rel-L2 values)
python -m pytest tests -m gpu -q --timeout 900 -rf -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/${P}_tests_full.log
# 2. bench lines: headline, poly_pc, reference arm
python bench.py --steps 20 --warmup 5 > gpurun_out/${P}_bench_n1.json 2> gpurun_out/${P}_bench_n1.err
python bench.py --workload poly_pc --steps 20 --warmup 5 --no-dsm --no-cpu-baseline > gpurun_out/${P}_bench_poly_pc.json 2> gpurun_out/${P}_bench_poly.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err
# 3. launch lists (per-launch duration + DRAM bytes + tensor activity) of one forward, one PC step at 64k Poly latents, one training step
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches_celeba_fwd_b1024.csv python tools/profile_forward.py celeba 1024 > /dev/null 2>&1
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches_poly_fwd_b64.csv python tools/profile_forward.py poly 64 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches_train_celeba_b256.csv python tools/profile_train.py celeba 256 > /dev/null 2>&1
# 4. ncu --set full of the dominant kernel (three launches: the K-long 16x16 layers: 3x3 512->256 + residual, 3x3 256->512 + GELU + statistics, 3x3 512->256 + residual + copy + statistics) and of the tensor-core depthwise kernel
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm_pair_kernel -s 6 -c 3 -o gpurun_out/${P}_pair --force-overwrite python tools/profile_forward.py celeba 1024 > /dev/null 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:dwconv7_mma -s 1 -c 1 -o gpurun_out/${P}_dwconv --force-overwrite python tools/profile_forward.py celeba 1024 > /dev/null 2>&1
# 5. per-layer GEMM timings and the depthwise A/B
python tools/bench_conv_shapes.py > gpurun_out/${P}_conv_shapes.log 2>&1
python tools/bench_dwconv.py > gpurun_out/${P}_dwconv.log 2>&1
SBM_DWCONV_MMA=0 python tools/bench_dwconv.py >> gpurun_out/${P}_dwconv.log 2>&1
python tools/epi_experiment.py > gpurun_out/${P}_epilogue_ab.log 2>&1
python tools/list_conv_variants.py > gpurun_out/${P}_conv_variants.log 2>&1
ls -la gpurun_out | tail -20
tail -c 600 gpurun_out/${P}_tests_full.log
head -c 300 gpurun_out/${P}_bench_n1.json
python tools/bench_openai.py > gpurun_out/${P}_bench_unetmodel.json 2>&1
python tools/bench_ae.py > gpurun_out/${P}_bench_res_ae.json 2>&1
python tools/time_train.py > gpurun_out/${P}_time_train.log 2>&1
for b in 128 256 512 1024; do python tools/time_forward.py $b 30; done > gpurun_out/${P}_forward_vs_batch.log 2>&1
