#!/bin/bash
# compute-sanitizer over a small subset of the GPU tests (SURVEY.md §5 "race detection"):
# memcheck on the scatter kernels (owner-computes tiles incl. the TMA store, global
# reductions), the binning kernel, the hand-written FFT passes (incl. the persistent z + y
# kernel with its inter-block flags) and the emulated slab decomposition; racecheck
# (shared-memory hazards) on the kernels that exchange data through shared memory.
# Logs: gpurun_out/sanitizer_{memcheck,racecheck}.txt
set -u
OUT=${1:-gpurun_out}
SEL_MEM='owner_computes and (TSC or PCS) or fused_zy or strided_fft or r2c_row or golden_double and (sim_tsc or survey_pcs) or streamed_host'
SEL_RACE='owner_computes and TSC and True or fused_zy or strided_fft or golden_double and sim_tsc_il'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitizer_memcheck.txt \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_MEM" > $OUT/sanitizer_memcheck_pytest.txt 2>&1
echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_pytest.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitizer_memcheck_slab.txt \
  python -m pytest tests/test_gpu_slab.py -m gpu -q -x -k "emulated_slabs_match_oracle" > $OUT/sanitizer_memcheck_slab_pytest.txt 2>&1
echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_slab_pytest.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/sanitizer_racecheck.txt \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_RACE" > $OUT/sanitizer_racecheck_pytest.txt 2>&1
echo "racecheck rc=$?" >> $OUT/sanitizer_racecheck_pytest.txt
for f in $OUT/sanitizer_*.txt; do echo "== $f"; tail -n 3 $f; done
