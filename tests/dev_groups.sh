#!/bin/bash
export CUDA_LAUNCH_BLOCKING=1
echo "=== ipa"; timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 200 -k "ipa_attention" 2>&1 | grep -E "^E   .*(Assert|Error)|passed|failed|^FAILED" | head -30
unset CUDA_LAUNCH_BLOCKING
echo "=== rest"; timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 200 -k "not ipa_attention" 2>&1 | grep -E "^E   .*(Assert|Error)|passed|failed|^FAILED" | head -30
echo "=== model"
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 2>&1 | grep -E "^E   .*(Assert|Error)|passed|failed|^FAILED" | head -30
