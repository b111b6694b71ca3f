#!/bin/bash
# compute-sanitizer over the kernel-level parity tests that wait on other CTAs, queue through shared
# memory or use mbarriers / bulk copies: the one-pass PAM scan (decoupled look-back, GUIDO_PAM_PASSES=1),
# the block join (warp pipelines, mbarriers, vote-ranked hit queue), the gathered seed index (+ the
# extras scattered from the bucket ends), the bit-parallel MMEJ kernel.
# ONE tool per gpurun call (B200_PROFILING.md): pass memcheck or racecheck.
# NOTE (round 2): on the build pool compute-sanitizer is closed ("runs under it have left GPUs needing a
# reset"; gpurun answers exit code 86 without running anything), so there is no sanitizer log in profiles/.
# The kernels carry their own checks instead: the gather counts buckets whose size disagrees with the
# scanned offsets (reported as an error by the host call), the PAM look-back has a bounded wait with an
# error flag, every parity test compares against the oracle / a second implementation.
#   gpurun --timeout 1500 -- 'tools/sanitize.sh memcheck'
tool=${1:-memcheck}
mkdir -p gpurun_out/logs
export GUIDO_PAM_PASSES=1
timeout 1400 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mmej.py -x -q -m gpu \
    -k "scan_matches_oracle and (NGG-kw0 or NAG) or join_variants or one_pass or gather_equals_scatter and 2-4 or csr_format or packed_hits or mmej_random or mmej_ragged or mmej_sites" \
    > gpurun_out/logs/r2_sanitize_$tool.log 2>&1
rc=$?
tail -15 gpurun_out/logs/r2_sanitize_$tool.log
echo "compute-sanitizer $tool exit code $rc"
exit $rc
