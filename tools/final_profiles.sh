# ncu evidence for the shipped kernels (run only after the plain commands have exited 0)
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ot_kernel -s 3 -c 1 -f -o gpurun_out/r1_ot_index python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pam > gpurun_out/ncu_ot_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pam_ -s 6 -c 2 -f -o gpurun_out/r1_pam python bench.py --workload pam --genome agam --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pam_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:idx_build -s 10 -c 9 -f -o gpurun_out/r1_idx python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pam > gpurun_out/ncu_idx_full.log 2>&1
ls -la gpurun_out | tail -12
