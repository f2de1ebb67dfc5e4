set -x
mkdir -p gpurun_out
# memcheck over the kernels that are new in round 2 (small inputs; the sanitizer slows kernels ~50x)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest -x -q -m gpu \
  tests/test_gpu_synth.py "tests/test_gpu_hash.py::test_hash_set_kats_on_gpu" "tests/test_gpu_hash.py::test_hash_membership_matches_the_oracle[21]" \
  "tests/test_gpu_set.py::test_set_built_from_a_stream_of_chunks_equals_the_one_shot_set[15]" \
  "tests/test_gpu_set.py::test_both_shapes_of_the_bucket_counting_kernel_agree[15]" > gpurun_out/sanitizer_r2_tests.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_r2_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_r2_smoke.log 2>&1
echo "memcheck smoke rc=$?"; tail -4 gpurun_out/sanitizer_r2_smoke.log
( timeout 900 python -m pytest tests/test_gpu_set.py tests/test_gpu_correct.py -m gpu -q -x ) > gpurun_out/r2l_tests.log 2>&1; tail -3 gpurun_out/r2l_tests.log
