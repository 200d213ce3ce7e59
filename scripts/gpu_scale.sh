# developer aid: the scaling bench at N GPUs of one box for cfg3 (headline) and cfg5 (BASELINE config 5)
# usage: gpurun --gpus N -- 'bash scripts/gpu_scale.sh N'
N=$1
mkdir -p gpurun_out
for w in cfg3 cfg5; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 --workload $w --no-cpu-baseline --no-shim --no-secondary > gpurun_out/r2_scale_${w}_n$N.json 2> gpurun_out/r2_scale_${w}_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --workload $w > gpurun_out/r2_scale_${w}_n$N.json 2> gpurun_out/r2_scale_${w}_n$N.err
  fi
  tail -c 300 gpurun_out/r2_scale_${w}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_scale_${w}_n$N.json").read().strip().splitlines()[-1])
    print("$w N=$N value %.0f it/s %.3f ms | e2e %.0f (%.3f ms) same-topology %.3f ms | %s | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["same_topology"]["ms_per_step"], {k: round(v, 4) for k, v in d["roofline"]["phase_ms_per_step"].items()}, d["numerics"]))
except Exception as e:
    print("$w N=$N: no line", e)
PY
done
if [ "$N" != "1" ]; then
  python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2 | tee gpurun_out/r2_mgpu_test_n$N.log
  # the worker's own report (rank agreement bit for bit, golden chi2, pre-sharded input), kept as evidence
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N tests/mgpu_worker.py 2>/dev/null | grep -v "^\[W\|^W[0-9]" | tee gpurun_out/r2_mgpu_worker_n$N.log | tail -12
fi
