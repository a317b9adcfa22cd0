# compute-sanitizer pass over a representative slice of the -m gpu suite (gpurun -- bash tools/gpu_sanitizer.sh).
# memcheck: sampler / DSM kernels, implicit-GEMM convolutions (single-CTA, CTA-pair, pixel-major, BN=64), attention
# (mma + small + backward), guidance, evaluators, the score nets end to end, smoke().  racecheck: the kernels with
# hand-rolled shared-memory hand-offs that are not mbarrier / TMA based (attention, depthwise, GroupNorm, pack / unpack).
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --launch-timeout 300 --log-file gpurun_out/r2_sanitizer_memcheck.log \
  python -m pytest tests/test_sampler_gpu.py tests/test_eval_samplers_gpu.py tests/test_guidance_gpu.py tests/test_unet_gpu.py \
  "tests/test_conv_igemm.py::test_conv_geometry" tests/test_backward_gpu.py -m gpu -q -x --timeout 1400 \
  -k "not full and not large_batch and not pair_lin and not pair_small2 and not pair64_lin and not tracks_reference" \
  > gpurun_out/r2_sanitizer_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck_pytest.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 --launch-timeout 300 --log-file gpurun_out/r2_sanitizer_racecheck.log \
  python -m pytest tests/test_backward_gpu.py tests/test_unet_gpu.py -m gpu -q -x --timeout 800 \
  -k "linear_attention or softmax_attention or dwconv or groupnorm or matches_oracle" \
  > gpurun_out/r2_sanitizer_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck_pytest.log
tail -5 gpurun_out/r2_sanitizer_memcheck_pytest.log gpurun_out/r2_sanitizer_racecheck_pytest.log
tail -5 gpurun_out/r2_sanitizer_memcheck.log gpurun_out/r2_sanitizer_racecheck.log
