# A/B of the data-parallel DSM step on N GPUs (gpurun --gpus 8 -- bash tools/dsm_dp_ab.sh 8): bucket size of the
# overlapped gradient all-reduce and the number of CTAs NCCL may take from the persistent GEMM kernels.
N=${1:-8}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $N --steps 10 --warmup 3 --dsm-only celeba 2>> gpurun_out/r2_dsm_ab.err | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'variant': '$name', 'ms_per_step': d['ms_per_step'], 'steps_per_s': d['value'], 'grad_allreduce': d['grad_allreduce']}))" >> gpurun_out/r2_dsm_ab_n$N.jsonl
}
run bucket64_bf16 SBM_DSM_BUCKET_MB=64
run bucket64_bf16_ctas4 SBM_DSM_BUCKET_MB=64 NCCL_MAX_CTAS=4
run bucket64_bf16_ctas8 SBM_DSM_BUCKET_MB=64 NCCL_MAX_CTAS=8
run single_bucket_bf16 SBM_DSM_BUCKET_MB=100000
run bucket256_bf16 SBM_DSM_BUCKET_MB=256
run single_bucket_fp32 SBM_DSM_BUCKET_MB=100000 SBM_DSM_COMM=fp32
cat gpurun_out/r2_dsm_ab_n$N.jsonl
