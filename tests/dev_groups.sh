#!/bin/bash
export CUDA_LAUNCH_BLOCKING=1
for g in "tensor_core" "conv5x5" "ipa_attention" "not (tensor_core or conv5x5 or ipa_attention)"; do
  echo "=== group: $g"
  timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 -k "$g" 2>&1 | grep -E "^E   .*(Assert|Error)|passed|failed|^FAILED" | head -30
done
unset CUDA_LAUNCH_BLOCKING
echo "=== model"
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 2>&1 | grep -E "^E   .*(Assert|Error)|passed|failed|^FAILED" | head -30
