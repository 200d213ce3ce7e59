# developer aid: bench lines only (no tests) at N GPUs: usage: gpurun --gpus N -- 'bash scripts/dev/scale_only.sh N [workloads]'
N=$1; shift
mkdir -p gpurun_out
for w in ${@:-cfg3 cfg5}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 3 --workload $w > gpurun_out/r2_scale_${w}_n$N.json 2> gpurun_out/r2_scale_${w}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_scale_${w}_n$N.json").read().strip().splitlines()[-1])
    print("$w N=$N value %.0f it/s %.3f ms | e2e %.0f (%.3f ms) same-topology %.3f ms | %s | bitwise %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["same_topology"]["ms_per_step"], {k: round(v, 4) for k, v in d["roofline"]["phase_ms_per_step"].items()}, d["numerics"].get("ranks_bitwise_equal")))
except Exception as e:
    print("$w N=$N: no line", e)
PY
done
