#!/bin/bash
# Run the GPU test tiers in separate processes so that one hung kernel cannot take the others down.
# usage: scripts/gpu_tests.sh [tier ...]   (default: all)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1 to=$2; shift 2
  echo "=== $name"
  timeout -k 10 "$to" python -m pytest -q -m gpu -p no:cacheprovider "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $? : $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ')"
}
tiers=${@:-"ops_base conv_fwd conv_dgrad conv_wgrad model input variants"}
for t in $tiers; do
  case $t in
    ops_base)   run ops_base 600 tests/test_ops_gpu.py -k "not tcgen05" ;;
    conv_fwd)   run conv_fwd 150 tests/test_ops_gpu.py -k "tcgen05 and conv_fwd" ;;
    conv_dgrad) run conv_dgrad 150 tests/test_ops_gpu.py -k "tcgen05 and conv_dgrad" ;;
    conv_wgrad) run conv_wgrad 150 tests/test_ops_gpu.py -k "tcgen05 and conv_wgrad" ;;
    model)      run model 900 tests/test_model_gpu.py ;;
    input)      run input 300 tests/test_tfrecord_gpu.py ;;
    variants)   run variants 900 tests/test_pix2pix_gpu.py tests/test_residual_gpu.py tests/test_bg_gpu.py tests/test_optimizers_gpu.py ;;
  esac
done
