#!/bin/bash
# ncu capture of the packed ribbon kernel (align) at the bench's launch shape + NTK checks after the wide-row generalisation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ntk_stages.py -x -q -m gpu > gpurun_out/r2w_pytest.log 2>&1; tail -2 gpurun_out/r2w_pytest.log
DYN_NTK_TRACE=1 timeout 300 python tools/ntk_k9_one.py 60 4 > gpurun_out/r2w_k9.log 2>&1; tail -5 gpurun_out/r2w_k9.log
CMD="python bench.py --config c2 --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/prof_r2w_align_plain.json 2> gpurun_out/prof_r2w_align_plain.err
ncu --set full --clock-control none --import-source on -k regex:k_ribbon -c 1 -f -o gpurun_out/prof_r2w_align $CMD > gpurun_out/prof_r2w_align_ncu.log 2>&1
tail -1 gpurun_out/prof_r2w_align_ncu.log
