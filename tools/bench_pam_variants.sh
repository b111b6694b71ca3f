#!/bin/bash
# PAM-scan bench once per library variant (guido_b200/lib/variants/*.so + the default build)
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1]));r=d['roofline']
print(sys.argv[2], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'hits', d['config']['hits'])" $1 $2; }
timeout 120 python bench.py --workload pam --genome agam --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p_default.json 2> gpurun_out/p_default.err && show gpurun_out/p_default.json default || tail -3 gpurun_out/p_default.err
for v in guido_b200/lib/variants/*.so; do
  n=$(basename $v .so)
  GUIDO_B200_LIB=$PWD/$v timeout 120 python bench.py --workload pam --genome agam --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p_$n.json 2> gpurun_out/p_$n.err && show gpurun_out/p_$n.json $n || tail -3 gpurun_out/p_$n.err
done
