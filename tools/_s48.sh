mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/s48_launches_poly_fwd_b32768.csv python tools/profile_forward.py poly 32768 > gpurun_out/s48_ncu.log 2>&1
timeout 200 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/s48_launches_celeba_fwd_b128.csv python tools/profile_forward.py celeba 128 >> gpurun_out/s48_ncu.log 2>&1
tail -3 gpurun_out/s48_ncu.log
