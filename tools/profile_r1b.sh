#!/bin/bash
# Round-1 (second session) evidence: bench line, ncu launch list of the same command, full captures of the kernels.
set -x
python bench.py --steps 5 --warmup 3 --also-fp32 > gpurun_out/bench_r1b_final.json 2> gpurun_out/bench_r1b_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1b_reference.json 2>> gpurun_out/bench_r1b_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l_r1b.log 2>&1
for p in fp64 fp32; do
  ncu --set full --clock-control none --import-source on -k regex:"tube_wide|src_kernel|pcm_kernel" -c 3 -o gpurun_out/prof_r1b_$p -f \
      python bench.py --utterances 4096 --seconds 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --precision $p > gpurun_out/ncu_r1b_$p.log 2>&1
done
# lane-per-section kernel on the single-utterance config (what small batches run)
TRM_TUBE_MAPPING=sections ncu --set full --clock-control none --import-source on -k regex:"tube_kernel" -c 1 -o gpurun_out/prof_r1b_sections_fp64 -f \
      python bench.py --utterances 256 --seconds 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --precision fp64 > gpurun_out/ncu_r1b_sections.log 2>&1
