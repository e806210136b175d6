set -x
python tools/prof_steady.py 4 2960 32 100 5 || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_main --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2_kmain_cfg4_steady_end python tools/prof_steady.py 4 2960 32 100 5 > gpurun_out/ncu_cfg4_end.log 2>&1
tail -3 gpurun_out/ncu_cfg4_end.log
for g in 2 4 8 16; do python bench.py --groups $g --no-cpu --no-extra --no-sweep --steps 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('groups', d['e2e']['groups'], 'e2e', d['e2e']['value'], 'spacing', d['e2e']['launch_spacing_us'], 'value', d['value'])"; done
