#!/bin/bash
# Final-state record: GPU parity suite, default bench line (both arms), launch list, full capture
# of the scoring kernels (PCA, GEMM-form SVM) and the run-based scan.
TAG=${TAG:-r1k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc $?"
cut -c1-300 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc $?"
cut -c1-300 gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --steps 1 --warmup 1 --fields 32 --pool 16 --chunk 16 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"scaler_pca|svm_rbf|label_scan|gate_kernel" -s 10 -c 7 \
    -o gpurun_out/prof_score_$TAG $CMD > gpurun_out/ncu_full_score_$TAG.log 2>&1
ls -la gpurun_out | tail -8
