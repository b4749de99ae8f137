#!/usr/bin/env bash
# compute-sanitizer over the GPU test-suite ("sanitizers in CI").  Run on a GPU box:
#
#   gpurun --timeout 2400 -- 'bash tools/sanitize.sh'            # all four tools (memcheck, initcheck, racecheck, synccheck), the default test selection
#   bash tools/sanitize.sh memcheck tests/test_compact.py         # one tool, explicit pytest arguments
#
# Logs go to gpurun_out/san_<tool>.log (full), the verdict lines of every run to gpurun_out/san_summary.txt; copy the summary to
# profiles/ when it is the evidence for a round.  Exit status: 0 iff every run finished with "0 errors" / no hazards AND all tests passed.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
SAN=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
TOOLS=${1:-all}
shift || true
# racecheck only sees shared-memory hazards: it gets the shared-memory-heavy code (tokeniser, planner / shared-node schedule, compact expander,
# prover FFTs).  memcheck / initcheck get the same plus the untrusted-record and ragged-size tests.  Sizes are small: the tools slow kernels 10-100x.
RACE_TESTS=${*:-"tests/test_wit_ingest.py tests/test_compact.py tests/test_columns.py tests/test_draw_retry.py tests/test_query_dedup.py tests/test_gpu_parity.py::test_stwo_shared_node_schedule_matches_oracle tests/test_gpu_parity.py::test_stwo_fixture_trace_bit_exact tests/test_gpu_prover.py::test_gpu_prover_matches_reference_prover"}
MEM_TESTS=${*:-"tests/test_wit_ingest.py tests/test_compact.py tests/test_columns.py tests/test_gpu_parity.py tests/test_config_space.py tests/test_draw_retry.py tests/test_query_dedup.py tests/test_s101_multiquery.py"}
# synccheck: the named barriers of the warp-specialised transcript kernel (every Stwo test runs it; the repeated-draw records take its uniform retry path)
SYNC_TESTS=${*:-"tests/test_draw_retry.py tests/test_gpu_parity.py::test_stwo_fixture_trace_bit_exact tests/test_query_dedup.py"}
rc_all=0
: > "$OUT/san_summary.txt"
run() { # tool, extra flags, tests
    local tool=$1 flags=$2 tests=$3 log="$OUT/san_$1.log"
    echo "== $tool: $tests" | tee -a "$OUT/san_summary.txt"
    # shellcheck disable=SC2086
    timeout 1500 "$SAN" --tool "$tool" $flags --error-exitcode 97 --target-processes all \
        python -m pytest $tests -m gpu -x -q -p no:cacheprovider > "$log" 2>&1
    local rc=$?
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" "$log" | tail -5 | tee -a "$OUT/san_summary.txt"
    echo "exit $rc" | tee -a "$OUT/san_summary.txt"
    [ $rc -ne 0 ] && rc_all=1
}
case "$TOOLS" in
    all|memcheck) run memcheck "--leak-check no" "$MEM_TESTS" ;;&
    all|initcheck) run initcheck "" "$MEM_TESTS" ;;&
    all|racecheck) run racecheck "--racecheck-report all" "$RACE_TESTS" ;;&
    all|synccheck) # synccheck flags a named barrier whose warps arrive from different program locations (tools/synccheck_probe.cu): this tool runs on a
                   # -DSSYM_SYNCCHECK build, which keeps the transcript kernel's hand-over barrier behind one location; the release build is restored after
        python stark-symphony_b200/build.py --synccheck > "$OUT/san_build_synccheck.log" 2>&1
        run synccheck "" "$SYNC_TESTS"
        python stark-symphony_b200/build.py --force > /dev/null 2>&1 ;;
esac
echo "overall: $([ $rc_all -eq 0 ] && echo CLEAN || echo FINDINGS)" | tee -a "$OUT/san_summary.txt"
exit $rc_all
