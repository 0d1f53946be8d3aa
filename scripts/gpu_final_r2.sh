#!/bin/bash
# Round-2 final GPU pass on 1 GPU: tests, smoke, bench line, launch list of the bench command, full ncu captures of the hot kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 600 gpurun_out/bench_r2d.json; tail -3 gpurun_out/bench_r2d.err
DFTATOM_OPTIONS="use_graph=0 stream_groups=1" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 0 --no-rn --no-batch --no-micro --no-cpu-baseline --no-parity > gpurun_out/launches_r2.log 2>&1
export DFT_OPTS="use_graph=0 stream_groups=1"
for k in search_rows match_win match_cta poisson_direct poisson_warm potential_energy density_update; do
  skip=12; if [ $k = match_cta ]; then skip=40; fi; if [ $k = poisson_warm ]; then skip=1; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_r2_$k python scripts/prof_c3.py > gpurun_out/prof_r2_$k.log 2>&1
  tail -1 gpurun_out/prof_r2_$k.log
done
