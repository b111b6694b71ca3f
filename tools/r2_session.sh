#!/bin/bash
# Round-2 GPU session: parity tests (log kept), the driver's default bench, then ncu evidence for the
# three shipped hot kernels.  Every ncu capture follows a plain run of the same command that exited 0.
#   gpurun --timeout 1500 -- 'bash tools/r2_session.sh [tests] [bench] [ncu]'
mkdir -p gpurun_out/logs
want() { [ $# -eq 0 ] && return 0; for a in $ARGS; do [ "$a" = "$1" ] && return 0; done; return 1; }
ARGS="${*:-tests bench ncu}"
B="python bench.py --no-cpu-baseline"
if want tests; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/logs/r2_gputests.log 2>&1; echo "gpu tests rc $?"; tail -4 gpurun_out/logs/r2_gputests.log
fi
if want bench; then
  timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r2_bench.err; cut -c1-3000 gpurun_out/r2_bench.json
  timeout 600 python bench.py --workload mmej > gpurun_out/r2_bench_mmej.json 2> gpurun_out/r2_bench_mmej.err; echo "mmej rc $?"; cut -c1-1500 gpurun_out/r2_bench_mmej.json
  timeout 600 python bench.py --workload locus > gpurun_out/r2_bench_locus.json 2> gpurun_out/r2_bench_locus.err; echo "locus rc $?"; cut -c1-1500 gpurun_out/r2_bench_locus.json
fi
if want ncu; then
  $B --no-pam --steps 2 --warmup 3 > gpurun_out/plain_ot.json 2> gpurun_out/plain_ot.err || { echo plain ot failed; exit 1; }
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv \
      $B --no-pam --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ot_join_blocks|idx_gather" -s 6 -c 2 -f \
      -o gpurun_out/r2_join $B --no-pam --steps 1 --warmup 3 > gpurun_out/ncu_join.log 2>&1
  $B --workload pam --genome agam --steps 4 --warmup 3 > gpurun_out/plain_pam.json 2> gpurun_out/plain_pam.err || { echo plain pam failed; exit 1; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_pam_launches.csv \
      $B --workload pam --genome agam --steps 2 --warmup 3 > gpurun_out/ncu_pam_list.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pam_onepass -s 6 -c 1 -f -o gpurun_out/r2_pam_onepass \
      $B --workload pam --genome agam --steps 4 --warmup 3 > gpurun_out/ncu_pam_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmej_fast_kernel -s 1 -c 1 -f -o gpurun_out/r2_mmej \
      $B --workload mmej --steps 1 --warmup 1 > gpurun_out/ncu_mmej.log 2>&1
  ls -la gpurun_out | tail -14
fi
