# one GPU session: parity tests, benches, then ncu captures of the same commands
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --workload pam --genome agam --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pam.json 2> gpurun_out/bench_pam.err; tail -2 gpurun_out/bench_pam.err; cut -c1-700 gpurun_out/bench_pam.json
timeout 300 python bench.py --workload offtarget --genome agam --guides 10000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ot_agam.json 2> gpurun_out/bench_ot_agam.err; tail -2 gpurun_out/bench_ot_agam.err; cut -c1-500 gpurun_out/bench_ot_agam.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ot_human.json 2> gpurun_out/bench_ot_human.err; tail -2 gpurun_out/bench_ot_human.err; cut -c1-1500 gpurun_out/bench_ot_human.json
if [ "$1" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_pam.csv python bench.py --workload pam --genome agam --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pam_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pam_scan -s 3 -c 1 -f -o gpurun_out/prof_pam python bench.py --workload pam --genome agam --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pam_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_ot.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ot_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ot_kernel -s 3 -c 1 -f -o gpurun_out/prof_ot python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ot_full.log 2>&1
ls -la gpurun_out
fi
