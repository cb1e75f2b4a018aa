#!/bin/bash
# compute-sanitizer over the round-2 additions: radiance cache / loss, fused Adam + densification, loss-tail kernels,
# SSIM, the fused step with the radiance term, the env tap cache. memcheck + synccheck (racecheck adds nothing new: none
# of these kernels stages data through shared memory across warps except the closest-hit stack, which is per thread).
tag=${1:-san2}
out=gpurun_out/$tag
mkdir -p $out
K="test_cache_matches_oracle or test_loss_and_gradients or test_fused_adam or densif or test_loss_terms or test_fused_ssim_matches or test_fused_step_carries or test_env_tap_cache or test_fused_tail_matches_torch"
for tool in ${2:-memcheck synccheck}; do
  timeout 1700 compute-sanitizer --tool $tool --report-api-errors no --error-exitcode 99 --print-limit 20 \
      python -m pytest tests/test_radiance_gpu.py tests/test_optim_gpu.py tests/test_loss_terms_gpu.py tests/test_ssim_gpu.py tests/test_fused_step_gpu.py tests/test_shading_gpu.py tests/test_fused_loss_gpu.py -m gpu -q -x -k "$K" > $out/$tool.log 2>&1
  echo "$tool exit $?" | tee -a $out/$tool.log
  grep -E "ERROR SUMMARY|passed|failed" $out/$tool.log | tail -4
done
