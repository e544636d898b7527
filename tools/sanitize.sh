#!/usr/bin/env bash
# compute-sanitizer pass over the GPU parity tests (run on a B200 box; SURVEY.md section 5 asks for it, the reference has
# nothing comparable).  memcheck on the kernel-level tests, racecheck on the shared-memory heavy kernels only -- the tools
# slow kernels down 10-100x, so the large-batch cases are deselected.
#   usage: gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
set -u
cd "$(dirname "$0")/.."
SEL='not large and not 100000 and not 190001 and not 65536 and not 150000'
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused.py -q -x -k "$SEL"
echo "memcheck exit code: $?"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 \
    python -m pytest tests/test_gpu_model.py -q -x -k "(edge_mlp or segment_pool) and not 70000 and $SEL"
echo "memcheck (edge MLP, both generations) exit code: $?"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 \
    python -m pytest tests/test_gpu_data_path.py tests/test_gpu_spectral.py tests/test_gpu_config3_exp.py -q -x -k "$SEL"
echo "memcheck (device collation, SpectralDesign, config 3) exit code: $?"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -x -k "(edge_mlp or gemm_tn or segment_pool) and not 70000 and $SEL"
echo "racecheck exit code: $?"
# last session of round 2: streaming act_bwd_y (+ gate weight gradients), 32-lane partial reduction, first-layer aggregate form,
# aligned_rows copy, dense SpectralDesign path
if [ "${SANITIZE_LATE:-1}" = "1" ]; then
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 \
    python -m pytest tests/test_gpu_z_added_late.py tests/test_gpu_modules.py -q -x -k "(act_bwd or aligned_rows or whole_layer or column_block) and not 40000 and $SEL"
echo "memcheck (late kernels) exit code: $?"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 \
    python -m pytest tests/test_gpu_z_added_late.py tests/test_gpu_fused.py -q -x -k "act_bwd and $SEL"
echo "racecheck (act_bwd_y kernels) exit code: $?"
fi
