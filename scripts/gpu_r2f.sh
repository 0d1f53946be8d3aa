#!/bin/bash
# Round-2 follow-up pass: shape tests after the host-side pivot tables, ncu captures of the warm Poisson kernels and of a tail-step search launch, reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scf.py -m gpu -x -q > gpurun_out/pytest_gpu_r2f.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2f.log
export DFT_OPTS="use_graph=0 stream_groups=1"
for k in poisson_direct poisson_warm; do
  skip=12; if [ $k = poisson_warm ]; then skip=1; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_r2_$k python scripts/prof_c3.py > gpurun_out/prof_r2_$k.log 2>&1
  tail -1 gpurun_out/prof_r2_$k.log
done
# tail step (3-10 atoms left): both shapes of the search kernel are launched every step, the 8-warp one does the work from step 32 on
timeout 300 ncu --set full --clock-control none --import-source on -k regex:search_rows -s 150 -c 2 -f -o gpurun_out/prof_r2_search_tail python scripts/prof_c3.py > gpurun_out/prof_r2_search_tail.log 2>&1
tail -1 gpurun_out/prof_r2_search_tail.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:match_cta -s 150 -c 1 -f -o gpurun_out/prof_r2_match_tail python scripts/prof_c3.py > gpurun_out/prof_r2_match_tail.log 2>&1
tail -1 gpurun_out/prof_r2_match_tail.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; tail -c 800 gpurun_out/bench_r2_ref.json
