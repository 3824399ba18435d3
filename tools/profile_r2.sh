#!/bin/bash
# Round-2 evidence: full ncu captures of the kernels of one step (4096 utterances x 0.5 s: same per-SM occupancy as the
# headline, 20 x shorter) in both precision modes, and the launch list of the default bench command.
set -x
for p in fp64 fp32; do
  ncu --set full --clock-control none --import-source on -k regex:"tube_wide|src_kernel|pcm_kernel" -c 3 -o gpurun_out/prof_r2_$p -f \
      python bench.py --utterances 4096 --seconds 0.5 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs --no-fast-mode --precision $p > gpurun_out/ncu_r2_$p.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-fast-mode > gpurun_out/ncu_l_r2.log 2>&1
