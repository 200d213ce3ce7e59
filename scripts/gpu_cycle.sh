# developer aid: parity + bench + solver trace in one gpurun call
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
for c in ${CLUSTERS:-4}; do
echo "== SSBA_SOLVE_CLUSTER=$c"
SSBA_SOLVE_CLUSTER=$c python bench.py --no-cpu-baseline --steps 50 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['roofline']['phase_ms_per_step'])"
done
SSBA_LIB=$PWD/ssvio_b200/lib/libssba_trace.so python scripts/solver_trace.py cfg3 > gpurun_out/trace_latest.log 2>&1; head -24 gpurun_out/trace_latest.log | tail -23
