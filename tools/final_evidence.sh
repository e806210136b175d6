#!/bin/bash
# End-of-round evidence run (one B200): the GPU suite, smoke(), both bench arms, the ncu launch list of the bench command and one
# `ncu --set full` capture of the bench kernel in the shipped shape.  Everything lands in gpurun_out/; the summaries that are
# judged are copied into profiles/ by hand (profiles/README.md names each file and its command).
#   gpurun --timeout 2400 -- 'bash tools/final_evidence.sh'
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -2 gpurun_out/r2_smoke_final.log
python bench.py --impl reference > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/r2_bench_reference_final.err
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 600 gpurun_out/r2_bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_simple --launch-skip 2 --launch-count 1 -f \
    -o gpurun_out/r2_ksimple_w8_final2 python tools/prof_run.py 4096 8 125 > gpurun_out/r2_ncu_full.log 2>&1
python examples/train_dqn.py --ticks 1000 > gpurun_out/r2_train_dqn_final.txt 2>&1; tail -3 gpurun_out/r2_train_dqn_final.txt
