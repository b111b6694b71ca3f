set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --workload pam --genome agam --steps 5 --warmup 3 > gpurun_out/bench_pam.json 2> gpurun_out/bench_pam.err; tail -3 gpurun_out/bench_pam.err; cat gpurun_out/bench_pam.json
timeout 600 python bench.py --workload offtarget --genome agam --guides 10000 --algo scan --steps 3 --warmup 3 > gpurun_out/bench_ot.json 2> gpurun_out/bench_ot.err; tail -3 gpurun_out/bench_ot.err; cat gpurun_out/bench_ot.json
