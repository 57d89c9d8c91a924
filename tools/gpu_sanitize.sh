#!/bin/bash
# compute-sanitizer over the GPU parity tests of the kernels written this round: bash tools/gpu_sanitize.sh <tag>
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
K='angle_3b or specialised_shapes or power_spectrum_shapes or variance_of_soap or tutorial or reference_data_all or variants_efv or determin or distance_2b_options or cutoff_skin or neighbour_list or si_two_descriptor or lammps or gap_xml_known or md_run_matches or reduce_hook'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck.log python -m pytest tests -m gpu -x -q -k "$K" > $OUT/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 $OUT/memcheck_pytest.log; tail -3 $OUT/memcheck.log
K2='angle_3b or specialised_shapes or variance_of_soap or variants_efv and (case0 or 55 or 117 or 23 or 113) or determin or distance_2b_options'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck.log python -m pytest tests -m gpu -x -q -k "$K2" > $OUT/racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
tail -3 $OUT/racecheck_pytest.log; tail -3 $OUT/racecheck.log
