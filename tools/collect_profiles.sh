#!/bin/bash
# Collects the round's measurement artefacts on a B200 box into gpurun_out/ (copied to profiles/ afterwards).
#   gpurun --timeout 1500 -- tools/collect_profiles.sh
set -u
O=gpurun_out
timeout 900 python bench.py > $O/r2_prof_bench_1gpu.json 2> $O/r2_prof_bench_1gpu.err
timeout 300 python bench.py --impl reference > $O/r2_prof_bench_reference.json 2>> $O/r2_prof_bench_1gpu.err
timeout 300 python bench.py --skip-secondary --no-pipeline > $O/r2_prof_bench_nopipeline.json 2>> $O/r2_prof_bench_1gpu.err
timeout 300 python tools/kernel_timeline.py > $O/r2_prof_timeline_block.log 2>&1
timeout 300 python tools/kernel_timeline.py --e2e > $O/r2_prof_timeline_e2e.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $O/r2_prof_launches_warm.csv \
    python tools/profile_step.py --updates 6 > /dev/null 2>&1
timeout 300 python tools/bench_kernels.py > $O/r2_prof_bench_kernels.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm_tc_kernel -s 6 -c 2 -f -o $O/prof_r2_gemm \
    python tools/profile_step.py --updates 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp3_forward_kernel -s 6 -c 3 -f -o $O/prof_r2_fused_fwd \
    python tools/profile_step.py --updates 4 > /dev/null 2>&1
(timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_cuda_kernels.py tests/test_cuda_update_parity.py -m gpu -q -x \
    -k "split or fused or gate or tail or pipelined or cross_call or rows or adam" 2>&1 | grep -v "Host Frame" | tail -25) > $O/r2_prof_memcheck.log
ls -la $O | tail -15
