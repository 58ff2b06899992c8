#!/bin/bash
# compute-sanitizer over the small-size parity suite: memcheck, initcheck, and racecheck on the shared-memory
# kernels (NTT passes, bitonic tail, scans, small-MSM trees).  Run on the GPU box: bash tools/sanitize.sh [outdir]
# Each tool's summary (ERROR SUMMARY line + any reports) goes to $OUT/sanitizer_<tool>.log.
OUT=${1:-gpurun_out}
mkdir -p $OUT
SEL='test_msm_kat or test_msm_adversarial_inputs or (test_msm_random_matches_oracle and (1000 or 4097)) or (test_msm_precomputed_tables and 11) or (test_msm_grouped_columns and 12) or (test_ntt_matches_oracle and (5 or 10 or 12 or 13)) or (test_coset_extension_matches_oracle and 9-11) or test_verify_accumulate_matches_oracle or test_fold_h_matches_oracle or test_g1_ops'
PLONK='(test_prover_writes_byte_identical_proofs and (my_circuit-6 or wide)) or test_verifier_glue_matches_oracle or test_kzg_setup_matches_oracle'
for tool in memcheck initcheck racecheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 20 --log-file $OUT/sanitizer_$tool.raw \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_plonk.py tests/test_gpu_params.py -m gpu -x -q -k "$SEL or $PLONK or test_params_file_round_trip" \
    > $OUT/sanitizer_$tool.pytest 2>&1
  echo "== $tool: pytest rc=$? ==" > $OUT/sanitizer_$tool.log
  tail -n 3 $OUT/sanitizer_$tool.pytest >> $OUT/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Uninitialized|hazard|Race reported" $OUT/sanitizer_$tool.raw | sort | uniq -c | head -40 >> $OUT/sanitizer_$tool.log
  head -c 20000 $OUT/sanitizer_$tool.raw > $OUT/sanitizer_$tool.head
done
cat $OUT/sanitizer_*.log
