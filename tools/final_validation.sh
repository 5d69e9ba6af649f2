#!/bin/bash
# Round-end validation on one B200 (run under gpurun from the repo root): the GPU test suite, smoke(), both
# bench arms, the ncu launch list of the bench command and one ncu --set full capture of the first three
# fused_round_kernel launches (claim probe, eval-only round 0, bind+eval round 1). Outputs -> gpurun_out/.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.log 2>&1
timeout 240 python bench.py > gpurun_out/final_ours.log 2>&1
grep -c "^{" gpurun_out/final_ref.log gpurun_out/final_ours.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01b.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-msm > gpurun_out/ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fused_round_kernel -c 3 -o /tmp/pf \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-msm > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/pf.ncu-rep --page raw --csv > gpurun_out/fused_r01b_raw.csv 2>/dev/null
ls -la gpurun_out/*.csv | tail -3
