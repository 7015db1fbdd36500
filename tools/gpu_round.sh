#!/bin/bash
# One GPU-box session: parity tests, microbenchmarks, bench lines, launch list, sanitizer.  Usage: tools/gpu_round.sh <tag> [parts]
# parts: any of  tests peak bench c1 ncu sani  (default: all)
set -u
TAG=${1:-r2a}; PARTS=${2:-"tests peak bench c1 ncu sani"}
O=gpurun_out; mkdir -p $O
has() { [[ " $PARTS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/${TAG}_smi.txt 2>&1
if has tests; then timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -5 $O/${TAG}_pytest_gpu.log; fi
if has peak; then ./build_tools/ffma_peak > $O/${TAG}_ffma_peak.txt 2>&1; tail -8 $O/${TAG}_ffma_peak.txt; fi
if has bench; then
  timeout 900 python bench.py --steps 100 --warmup 10 > $O/${TAG}_bench_C2.json 2> $O/${TAG}_bench_C2.err; tail -c 600 $O/${TAG}_bench_C2.err
  python - <<PY
import json
try:
    d = json.loads(open('$O/${TAG}_bench_C2.json').read().strip().splitlines()[-1])
    print('C2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'unchanged', d['e2e_unchanged_ppo']['ms_per_step'])
    print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], 'hbm', d['roofline']['hbm']['frac'])
    for k, v in d['per_config'].items(): print(k, {x: v.get(x) for x in ('ms_per_step', 'value', 'workspace_gb', 'error')}, v.get('e2e', {}).get('ms_per_step'))
    for t in d['top_kernels']: print(t)
    print('cpu', d.get('cpu_baseline'))
except Exception as e: print('bench parse failed', e)
PY
fi
if has c1; then timeout 600 python bench.py --workload C1 --steps 50 --warmup 5 > $O/${TAG}_bench_C1.json 2> $O/${TAG}_bench_C1.err; tail -c 300 $O/${TAG}_bench_C1.err; cut -c1-600 $O/${TAG}_bench_C1.json; fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/${TAG}_launches_bench_C2.csv \
      python bench.py --steps 4 --warmup 3 --no-per-config --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
  python tools/summarize_launches.py $O/${TAG}_launches_bench_C2.csv > $O/${TAG}_launches_bench_C2.txt 2>&1; head -30 $O/${TAG}_launches_bench_C2.txt
fi
if has sani; then
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_step.py C2 24 > $O/${TAG}_memcheck_C2.log 2>&1; echo memcheck rc $?; tail -4 $O/${TAG}_memcheck_C2.log
  timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python tools/sanitize_step.py C2 12 > $O/${TAG}_racecheck_C2.log 2>&1; echo racecheck rc $?; tail -4 $O/${TAG}_racecheck_C2.log
fi
ls -la $O | tail -15
