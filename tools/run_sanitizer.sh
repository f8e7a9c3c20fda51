#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck) over the hand-written tcgen05 / TMA / mbarrier kernels, through small GPU tests.
# Usage (on the GPU box): bash tools/run_sanitizer.sh [outdir]   -- logs go to <outdir>/sanitizer_<tool>.log
out=${1:-gpurun_out}
mkdir -p "$out"
TESTS="tests/test_gpu_conv_umma.py::test_every_conv_umma_vs_simt_and_cpu[shape1] tests/test_gpu_output_parity.py::test_every_conv_tf32_vs_exact_tf32_arithmetic[shape1] tests/test_gpu_conv_umma.py::test_fused_basic_block_vs_separate_convs_and_cpu[shape2] tests/test_gpu_vitpose.py::test_gemm_umma[2100-1152-384-0-False-True] tests/test_gpu_vitpose.py::test_gemm_umma[300-384-384-0-True-False] tests/test_gpu_vitpose.py::test_gemm_umma[130-128-64-2-False-True] tests/test_gpu_vitpose.py::test_attention_umma[2-128] tests/test_gpu_vitpose.py::test_attention_umma[3-60] tests/test_gpu_parity.py::test_uplift_bf16_tensor_core_bound"
for tool in memcheck synccheck racecheck; do
  log="$out/sanitizer_$tool.log"
  echo "== compute-sanitizer --tool $tool ==" > "$log"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python -m pytest $TESTS -x -q -m gpu -p no:cacheprovider >> "$log" 2>&1
  echo "exit code $?" >> "$log"
  tail -4 "$log"
done
