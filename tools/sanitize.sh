#!/bin/bash
# compute-sanitizer over the kernels with shared-memory hand-offs, mbarriers, TMEM and TMA:
# memcheck + racecheck + synccheck on smoke() and on the small-batch / ragged-tile parity tests of the fused tick,
# the tcgen05 predictors, the policy kernels and the one-lane (TMA tile) tick.  Run on a GPU box:
#     bash tools/sanitize.sh [out_dir]            (writes one summary per tool; exit code 0 = no hazards reported)
set -u
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
PY=${PYTHON:-python}
TESTS=${SAN_TESTS:-"tests/test_gpu_parity.py::test_rollout_fused_kernel_equals_per_tick_launches tests/test_gpu_wide.py::test_tp_ring_window_equals_shifted_window tests/test_gpu_parity.py::test_fused_tick_predictor_kernel_equals_two_launches tests/test_gpu_parity.py::test_ragged_batch_and_done_tick tests/test_gpu_wide.py::test_wide_equals_narrow_bit_for_bit tests/test_policy.py tests/test_gpu_parity.py::test_fused_predictor_matches_torch_lstm"}
KEXPR='not 9500 and not 4096 and not 4100 and not 2052'      # the large-batch parametrisations only repeat the small ones (the tools slow kernels 10-100x)
rc=0
for tool in memcheck racecheck synccheck; do
    log="$OUT/${tool}.log"
    echo "== $tool: smoke()" > "$log"
    if [ "${SAN_NO_SMOKE:-0}" != "1" ]; then
        compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 $PY __graft_entry__.py smoke >> "$log" 2>&1 || rc=1
    fi
    echo "== $tool: pytest" >> "$log"
    compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 $PY -m pytest -x -q -m gpu $TESTS -k "$KEXPR" >> "$log" 2>&1 || rc=1
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|passed|failed|smoke ok" "$log" | sort | uniq -c > "$OUT/${tool}.summary"
    cat "$OUT/${tool}.summary"
done
exit $rc
