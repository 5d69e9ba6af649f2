#!/bin/bash
# compute-sanitizer over the GPU suite (the CUDA analogue of the reference's lint / clippy discipline).
# The resident kernel waits for host commands, which a serialising tool tolerates only because the host never blocks on
# it - JB_FORCE_RESIDENT=1 keeps it enabled under the tool (by default the library falls back to one launch per round
# when it detects CUDA injection). Outputs -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
export JB_RESIDENT_TIMEOUT_S=200   # a tool can hold the host for many seconds while it patches a module
SEL='not 2pow22 and not 2pow18 and not 2pow16 and not baseline and not resident_equals_launch and not scheduler_batches and not lost_resident'
FILES="tests/test_gpu_field.py tests/test_gpu_bind.py tests/test_gpu_eq.py tests/test_gpu_sumcheck.py tests/test_gpu_resident.py tests/test_gpu_spliteq.py tests/test_gpu_compact.py tests/test_gpu_batch_add.py"
JB_FORCE_RESIDENT=1 timeout 900 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 7 \
    python -m pytest $FILES -q -x -k "$SEL" --timeout 600 > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
JB_FORCE_RESIDENT=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 \
    python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_resident.py tests/test_gpu_eq.py tests/test_gpu_batch_add.py -q -x \
    -k "$SEL and not lockstep_125 and not large_one_hot" --timeout 600 > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
JB_FORCE_RESIDENT=1 timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 \
    python -m pytest tests/test_gpu_resident.py tests/test_gpu_sumcheck.py -q -x -k "$SEL" --timeout 600 > gpurun_out/sanitizer_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/sanitizer_synccheck.log
for f in memcheck racecheck synccheck; do tail -n 4 gpurun_out/sanitizer_$f.log; done
