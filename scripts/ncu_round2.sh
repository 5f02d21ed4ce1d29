#!/bin/bash
# ncu evidence of round 2 (run on the GPU box after the plain commands have exited 0 without ncu; results under gpurun_out/, the
# summaries are written to profiles/ by scripts/summarize_profiles.py).
#   1. --set full of the operator kernels at C3 and C5 size (k_layer_pre, k_layer_forward, k_vjp_phase_a, k_vjp_phase_b)
#   2. launch list (gpu__time_duration) of one C3 training step of bench.py
#   3. --set full of the two history kernels deep inside a C3 backward solve (DRAM traffic vs algorithmic bytes)
set -x
OUT=gpurun_out
OPS='regex:k_layer_pre|k_layer_forward|k_vjp_phase_a|k_vjp_phase_b'
python scripts/prof_operator.py c3 2 > $OUT/r02n_op_c3_plain.log 2>&1 || exit 1
python scripts/prof_operator.py c5 2 > $OUT/r02n_op_c5_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k "$OPS" --launch-skip 4 -c 8 -f -o $OUT/r02n_op_c3 python scripts/prof_operator.py c3 2 > $OUT/r02n_op_c3_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k "$OPS" --launch-skip 4 -c 8 -f -o $OUT/r02n_op_c5 python scripts/prof_operator.py c5 2 > $OUT/r02n_op_c5_ncu.log 2>&1
PSI_BENCH_MIN_WARMUP=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-blocks > $OUT/r02n_bench_plain.json 2> $OUT/r02n_bench_plain.err || exit 1
PSI_BENCH_MIN_WARMUP=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02n_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-blocks > $OUT/r02n_launches.log 2>&1
PSI_BENCH_MIN_WARMUP=1 ncu --set full --clock-control none --import-source on -k 'regex:k_qn_dots_tma|k_qn_axpy_tma' --launch-skip 900 -c 4 -f -o $OUT/r02n_hist python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-blocks > $OUT/r02n_hist_ncu.log 2>&1
ls -la $OUT | grep r02n
