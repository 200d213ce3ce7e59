# developer aid: compute-sanitizer evidence (memcheck, racecheck, synccheck) over the kernels of the LM trial for the
# three cluster shapes of the reduced solver (1, 4, 8 CTAs), the level solver, the pose-only and pose-graph paths
# usage: gpurun -- 'bash scripts/gpu_sanitize.sh'
mkdir -p gpurun_out
out=gpurun_out/r2_sanitizer.log
: > $out
run() { # tool, label, command...
  tool=$1; label=$2; shift 2
  echo "== $tool | $label" | tee -a $out
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|hazard|Invalid|chi2|========= Error|Barrier error" | head -12 | tee -a $out
}
for tool in memcheck racecheck synccheck; do
  run $tool "small (1 CTA), 3 iterations" python scripts/run_case.py small 3
  run $tool "cfg2 (4-CTA cluster), 3 iterations" python scripts/run_case.py cfg2 3
  run $tool "cfg3 (8-CTA cluster), 2 iterations" python scripts/run_case.py cfg3 2
done
SSBA_SOLVER=level run memcheck "cfg2, level solver (k_reduced_solve)" python scripts/run_case.py cfg2 3
SSBA_SOLVER=level run racecheck "cfg2, level solver (k_reduced_solve)" python scripts/run_case.py cfg2 3
run memcheck "pose-only + pose-graph tests" python -m pytest tests/test_pose_only.py tests/test_pose_graph.py -m gpu -q -x
echo done | tee -a $out
