#!/bin/bash
# Final single-GPU evidence (second half of round 2): GPU test suite, the default bench line, the reference arm, config 4 on
# one GPU, the ncu launch list of the bench command and one full ncu capture of the step kernel.  Logs: gpurun_out/r2z_*.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== GPU test suite"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2z_pytest_gpu.log
echo "== bench (default flags of the driver)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 300 gpurun_out/r2z_bench.json; tail -2 gpurun_out/r2z_bench.err
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; tail -c 300 gpurun_out/r2z_bench_reference.json
echo "== config 4 on one GPU"; timeout 600 python bench.py --config amazon500 --steps 30 --warmup 5 --no-init-a --no-cpu > gpurun_out/r2z_bench_amazon500.json 2> gpurun_out/r2z_bench_amazon500.err
echo "== config 2"; timeout 600 python bench.py --config enron50 --steps 30 --warmup 5 --no-init-a --no-cpu --no-traffic > gpurun_out/r2z_bench_enron50.json 2> gpurun_out/r2z_bench_enron50.err
python - <<'PY'
import json
for c in ['', '_amazon500', '_enron50']:
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2z_bench{c}.json') if l.startswith('{')][-1])
        ls=d.get('line_search') or {}
        print(c or 'amazon200', 'ms/step %.4f kernel %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']), 'searched', ls.get('nodes_line_searched'), 'of', ls.get('nodes_asked'), 'exhaustive kernel', (ls.get('exhaustive') or {}).get('step_kernel_ms'), 'conv', ls.get('run_to_convergence'))
    except Exception as e:
        print(c, 'FAILED', e)
PY
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-init-a --no-traffic > gpurun_out/r2z_launches_bench.log 2>&1; grep -c tile_step_kernel gpurun_out/r2z_launches.csv
echo "== ncu --set full of the step kernel (steady state)"; bash tools/r2_gpu_ncu.sh r2z_prof_tile 14
