#!/bin/bash
# batch-size sweep of the kernels (2 s utterances), PCIe probe, pipeline trace
set -x
python tools/pcie_probe.py > gpurun_out/pcie.log 2>&1
for n in 512 1024 2048 4096 8192 16384 32768; do
  python bench.py --utterances $n --seconds 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --also-fp32 > gpurun_out/sweep_$n.json 2> gpurun_out/sweep_$n.err
done
TRM_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/trace64.json 2> gpurun_out/trace64.err
TRM_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/trace32.json 2> gpurun_out/trace32.err
