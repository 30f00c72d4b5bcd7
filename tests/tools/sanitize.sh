#!/bin/bash
# compute-sanitizer passes over the 720p golden parity tests (every post op shape, every effect entry point):
# memcheck (out-of-bounds / misaligned accesses) and racecheck (shared-memory hazards in the blur rings, the landscape's
# transposition buffer, the LUT staging).  Logs go to gpurun_out/; summaries are copied to profiles/.
#   tests/tools/sanitize.sh [per-tool timeout in seconds]
T=${1:-900}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/sanitizer_summary.txt
TESTS="tests/test_gpu_post.py::test_post_case_matches_golden tests/test_gpu_effects.py::test_every_golden_effect_case tests/test_gpu_post.py::test_blend_chain_equals_sequential_blends tests/test_gpu_post.py::test_old_blur_every_kernel_size_in_place tests/test_gpu_beam_tail.py::test_accumulated_step_of_every_tail_length[1280] tests/test_gpu_beam_tail.py::test_every_tail_length[13153440-0.0-1280] tests/test_gpu_effects.py::test_effect_vs_live_reference_4k[landscape@500] tests/test_gpu_effects.py::test_effect_vs_live_reference_4k[ball@2060]"
for tool in memcheck racecheck; do
	timeout $T compute-sanitizer --tool $tool --error-exitcode 86 --log-file $OUT/sanitizer_$tool.log \
		python -m pytest $TESTS -m gpu -q -x -p no:cacheprovider > $OUT/sanitizer_${tool}_pytest.log 2>&1
	echo "$tool exit=$?" >> $OUT/sanitizer_summary.txt
	tail -3 $OUT/sanitizer_${tool}_pytest.log >> $OUT/sanitizer_summary.txt
	grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/sanitizer_$tool.log | tail -2 >> $OUT/sanitizer_summary.txt
done
cat $OUT/sanitizer_summary.txt
