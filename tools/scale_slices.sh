# N = 8 (and 4): the step with 1, 2 and 4 key slices per rank and with the peer-copy / NCCL merge
mkdir -p gpurun_out
for n in ${NS:-8 4}; do for cfg in ${CFGS:-"1 peer" "2 peer" "2 nccl"}; do
  set -- $cfg
  GUIDO_BENCH_SLICES=$1 GUIDO_BENCH_MERGE=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n*10+$1)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/sl_${n}_$1_$2.json 2> gpurun_out/sl_${n}_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sl_${n}_$1_$2.json").read().strip().splitlines()[-1])
    print($n, "slices", $1, "$2", "step", round(d["ms_per_step"],3), "join", round(d["roofline"]["kernel_ms"],3), "e2e", round(d["e2e"]["ms_per_step"],2), d["config"]["hits"], d["config"]["hits_checksum"], d["config"]["merge"])
except Exception as e:
    print($n, "$cfg", "failed", e)
PY
done; done
