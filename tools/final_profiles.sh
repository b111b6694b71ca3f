# ncu evidence for the shipped kernels (run each capture only after the plain command has exited 0).
# One GPU; ~3 GPU-minutes in total.  Outputs land in gpurun_out/; summarise with tools/ncu_brief.py
# and copy what should be judged into profiles/.
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pam > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err || exit 1
# launch list of the default step (index build next to the join, two join streams)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pam > gpurun_out/ncu_list.log 2>&1
# the join kernel alone (GUIDO_OT_OVERLAP=0: one launch over the whole table)
GUIDO_OT_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:ot_join_sliced -s 3 -c 1 -f \
    -o gpurun_out/r_join python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pam > gpurun_out/ncu_join.log 2>&1
# PAM scan: launch list + the one-pass kernel
python bench.py --workload pam --genome agam --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_pam.json 2> gpurun_out/plain_pam.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r_pam_launches.csv \
    python bench.py --workload pam --genome agam --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pam_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pam_onepass -s 6 -c 1 -f -o gpurun_out/r_pam_onepass \
    python bench.py --workload pam --genome agam --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pam_full.log 2>&1
# index build and hit sort kernels, if they are the subject
# ncu --set full --clock-control none --import-source on -k regex:"idx_build|ball_count|seg_rank_sort|hit_scatter" -s 10 -c 8 -f -o gpurun_out/r_idx python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pam
ls -la gpurun_out | tail -12
