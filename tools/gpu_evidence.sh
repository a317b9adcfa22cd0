# Round-end evidence run on the GPU box (gpurun -- bash tools/gpu_evidence.sh): full GPU suite with printed
# measurements, bench lines, ncu launch lists and --set full captures; everything lands in gpurun_out/r2_i_*.
set -x
mkdir -p gpurun_out
# 1. whole GPU suite with the printed measurements (drift curves, rel-L2 values)
python -m pytest tests -m gpu -q --timeout 900 -rf -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r2_i_tests_full.log
# 2. bench lines: headline, poly_pc
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_i_bench_n1.json 2> gpurun_out/r2_i_bench_n1.err
python bench.py --workload poly_pc --steps 20 --warmup 5 --no-dsm --no-cpu-baseline > gpurun_out/r2_i_bench_poly_pc.json 2> gpurun_out/r2_i_bench_poly.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_i_bench_reference.json 2> gpurun_out/r2_i_bench_reference.err
# 3. launch lists (per-launch duration + DRAM bytes + tensor activity) of one forward, one PC step at 64k Poly latents, one training step
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_i_launches_celeba_fwd_b1024.csv python tools/profile_forward.py celeba 1024 > /dev/null 2>&1
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_i_launches_sampler_64k.csv python tools/profile_sampler.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_i_launches_train_celeba_b256.csv python tools/profile_train.py celeba 256 > /dev/null 2>&1
# 4. ncu --set full of the dominant kernel (two launches) and of the new attention kernel
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm_pair_kernel -s 20 -c 2 -o gpurun_out/r2_i_pair --force-overwrite python tools/profile_forward.py celeba 1024 > /dev/null 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:linear_attn_mma -c 1 -o gpurun_out/r2_i_attn --force-overwrite python tools/profile_forward.py celeba 1024 > /dev/null 2>&1
ls -la gpurun_out | tail -20
tail -c 600 gpurun_out/r2_i_tests_full.log
head -c 300 gpurun_out/r2_i_bench_n1.json
