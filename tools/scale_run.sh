# 1 -> 8 GPU scaling of the headline bench on one box
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-pam > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
done
for n in 1 2 4 8; do python - <<PY
import json
d=json.loads(open("gpurun_out/scale_$n.json").read().strip().splitlines()[-1])
print($n, round(d["ms_per_step"],2), "ms", round(d["value"]/1e6,2), "M guides/s", "kernel", round(d["roofline"]["kernel_ms"],2), "e2e", round(d["e2e"]["ms_per_step"],1), "hits", d["config"]["hits"])
PY
done
