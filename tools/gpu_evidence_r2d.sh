# Final evidence run of round 2 (gpurun -- bash tools/gpu_evidence_r2d.sh): GPU suite with printed measurements, bench
# lines, ncu launch lists and --set full captures of the dominant kernel and of the split-K pair; files gpurun_out/r2d_*.
P=r2d
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -rf -s 2>&1 | grep -v "^$" | tail -170 > gpurun_out/${P}_gpu_tests_printed_measurements.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${P}_bench_n1.json 2> gpurun_out/${P}_bench_n1.err
python bench.py --workload poly_pc --steps 20 --warmup 5 --no-dsm --no-cpu-baseline > gpurun_out/${P}_bench_poly_pc.json 2> gpurun_out/${P}_bench_poly.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches_celeba_fwd_b1024.csv python tools/profile_forward.py celeba 1024 > /dev/null 2>&1
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches_celeba_fwd_b128.csv python tools/profile_forward.py celeba 128 > /dev/null 2>&1
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches_poly_fwd_b64.csv python tools/profile_forward.py poly 64 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches_train_celeba_b256.csv python tools/profile_train.py celeba 256 > /dev/null 2>&1
# --set full: the three K-long 16x16 layers of the dominant kernel (batch 1024); one split-K GEMM pass + its epilogue kernel (batch 128)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm_pair_kernel -s 6 -c 3 -o gpurun_out/${P}_pair --force-overwrite python tools/profile_forward.py celeba 1024 > /dev/null 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:"131072|conv_splitk_epilogue" -s 4 -c 2 -o gpurun_out/${P}_splitk --force-overwrite python tools/profile_forward.py celeba 128 > /dev/null 2>&1
python tools/list_conv_variants.py > gpurun_out/${P}_conv_variants.log 2>&1
python tools/list_conv_variants.py 128 > gpurun_out/${P}_conv_variants_b128.log 2>&1
python tools/bench_openai.py > gpurun_out/${P}_bench_unetmodel.json 2>&1
python tools/time_train.py > gpurun_out/${P}_time_train.log 2>&1
ls -la gpurun_out | tail -24
tail -c 400 gpurun_out/${P}_gpu_tests_printed_measurements.log
head -c 300 gpurun_out/${P}_bench_n1.json
