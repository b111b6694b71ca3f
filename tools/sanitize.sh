#!/bin/bash
# compute-sanitizer over the smoke workload and the kernel-level parity tests that wait on other
# CTAs or queue through shared memory: the one-pass PAM scan (decoupled look-back, GUIDO_PAM_PASSES=1),
# the block join (warp pipelines, mbarriers, shared-memory hit queue) with the index build overlapped.
# ONE tool per gpurun call (B200_PROFILING.md): pass memcheck or racecheck.
#   gpurun --timeout 900 -- 'tools/sanitize.sh memcheck'
tool=${1:-memcheck}
mkdir -p gpurun_out/logs
export GUIDO_PAM_PASSES=1
timeout 800 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py -x -q -m gpu \
    -k "scan_matches_oracle and (NGG-kw0 or NAG) or join_variants or one_pass or block_join_full_passes and 4 or packed" \
    > gpurun_out/logs/r2_sanitize_$tool.log 2>&1
rc=$?
tail -15 gpurun_out/logs/r2_sanitize_$tool.log
echo "compute-sanitizer $tool exit code $rc"
exit $rc
