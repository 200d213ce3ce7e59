# developer aid (round 2): one gpurun call = microbenchmarks + parity + bench of both reduced solvers
# usage: gpurun -- 'bash scripts/gpu_r2_cycle.sh [what ...]'   what: dmma parity bench cfg5 all
mkdir -p gpurun_out
WHAT="${@:-all}"
has() { case " $WHAT " in *" $1 "*|*" all "*) return 0;; esac; return 1; }
summ='import json,sys
for l in sys.stdin:
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l); r=d.get("roofline") or {}
    print("value %.0f it/s  %.3f ms/step | e2e %.0f it/s %.3f ms | phases %s | chi2 %.8f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: round(v,4) for k,v in (r.get("phase_ms_per_step") or {}).items()}, d["config"]["chi2_robust_final"]))'
if has dmma; then
  (cd scripts/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu && timeout 120 ./dmma) > gpurun_out/r2_dmma.txt 2>&1
  tail -40 gpurun_out/r2_dmma.txt
fi
if has sanitize; then
  for cfg in ${SAN_CFGS:-small cfg2}; do
    echo "== compute-sanitizer memcheck $cfg"
    timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/run_case.py $cfg 3 2>&1 | grep -v "^=========     at\|^=========         in\|Host Frame" | tail -25 | tee gpurun_out/r2_memcheck_$cfg.log
  done
fi
if has parity; then
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r2_pytest_gpu.log
fi
if has bench; then
  for solver in tree level; do
    echo "== cfg3 solver=$solver"
    SSBA_SOLVER=$solver timeout 600 python bench.py --no-cpu-baseline --steps 50 2> gpurun_out/r2_bench_cfg3_$solver.err | tee gpurun_out/r2_bench_cfg3_$solver.json | python -c "$summ"
    tail -3 gpurun_out/r2_bench_cfg3_$solver.err
  done
fi
if has cfg5; then
  for solver in tree level; do
    echo "== cfg5 solver=$solver"
    SSBA_SOLVER=$solver timeout 900 python bench.py --workload cfg5 --no-cpu-baseline --steps 20 2> gpurun_out/r2_bench_cfg5_$solver.err | tee gpurun_out/r2_bench_cfg5_$solver.json | python -c "$summ"
    tail -3 gpurun_out/r2_bench_cfg5_$solver.err
  done
fi
