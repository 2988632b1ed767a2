#!/usr/bin/env bash
# compute-sanitizer passes over the GPU tests (run on a B200: `gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'`).
#   memcheck  : every kernel, on the small-shape tests (the sanitizer slows kernels 10-100x; the W >= 1e5 cases are deselected)
#   racecheck : shared-memory hazards of the non-cooperative kernels (multi-pair engine, dense / ami scoring, k-means update)
#   synccheck : barrier usage of the persistent cooperative kernels
# Exit status: non-zero if any pass reports an error.
set -uo pipefail
cd "$(dirname "$0")/.."
SAN=${SAN:-compute-sanitizer}
SMALL='not 100003 and not 100_003 and not 300000 and not 1000003 and not 50000 and not 20000-10 and not pipeline and not run_sh'
status=0
run() { echo "=== $*"; timeout ${SAN_TIMEOUT:-600} "$@"; rc=$?; echo "=== exit $rc"; [ $rc -eq 0 ] || status=1; }
run $SAN --tool memcheck --error-exitcode 1 python -m pytest -q -x -m gpu -k "$SMALL" \
    tests/test_zz_mi_pairs_gpu.py tests/test_zz_ami_gpu.py tests/test_dense_mi_gpu.py tests/test_batch_mi_gpu.py
run $SAN --tool memcheck --error-exitcode 1 python -m pytest -q -x -m gpu -k "ragged or strided or warmup or update_is_bit_exact or split_update or (tensor_equals_exact and (128-64-16 or 129-64-256 or 300-88-13))" \
    tests/test_kmeans_gpu.py
run $SAN --tool memcheck --error-exitcode 1 python -m pytest -q -x -m gpu -k "(matches_c_oracle and 257-3-256) or add_samples or errors or mixed or uniform_ids" tests/test_mi_gpu.py
run $SAN --tool memcheck --error-exitcode 1 python -m pytest -q -x -m gpu -k "byte_stream_variants and (60000-1024-400-6-0-1 or 120000-300-700-5-3-1)" tests/test_mi_gpu.py
# byte-stream layout kernels (block sort, warp-per-block arrangement with aliased scratch) and the scan's shared tables
run $SAN --tool racecheck --error-exitcode 1 python -m pytest -q -x -m gpu -k "byte_stream_variants and 60000-1024-400-6-0-1" tests/test_mi_gpu.py
# km_update_bulk_kernel is excluded from racecheck: its ring is written by cp.async.bulk (async proxy) and read by the
# consumer warp under full/empty mbarriers, an ordering racecheck does not model -- it reports the bulk-copy write
# against every consumer read (profiles/r02_sanitize_b.log keeps that report); memcheck and the bit-exact tests cover it.
run $SAN --tool racecheck --kernel-name-exclude kernel_substring=km_update_bulk_kernel --error-exitcode 1 python -m pytest -q -x -m gpu -k "update_is_bit_exact and (4096-128-8 or 65-88-17 or 4096-130-8 or 1025-16-2600)" tests/test_kmeans_gpu.py
run $SAN --tool racecheck --error-exitcode 1 python -m pytest -q -x -m gpu -k "reference_bits or one_pair or subset_of_columns" \
    tests/test_zz_mi_pairs_gpu.py tests/test_zz_ami_gpu.py
run $SAN --tool synccheck --error-exitcode 1 python -m pytest -q -x -m gpu -k "uniform_ids or mixed" tests/test_mi_gpu.py
exit $status
