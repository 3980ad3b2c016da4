#!/bin/bash
# Final single-GPU evidence of the round: GPU test suite, the default bench line, the reference arm, the ncu launch
# list of the bench command and one full ncu capture of the step kernel.  Logs under gpurun_out/r2z_*.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== GPU test suite"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2z_pytest_gpu.log
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 300 gpurun_out/r2z_bench.json; tail -2 gpurun_out/r2z_bench.err
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; tail -c 400 gpurun_out/r2z_bench_reference.json
echo "== other configs (1 GPU)"
for cfg in enron50 amazon500; do
  timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-init-a > gpurun_out/r2z_bench_$cfg.json 2> gpurun_out/r2z_bench_$cfg.err
done
BIGCLAM_TILE_EDGES=0 timeout 600 python bench.py --config amazon500 --steps 30 --warmup 5 --no-init-a --no-cpu --no-traffic > gpurun_out/r2z_bench_amazon500_notiles.json 2> gpurun_out/r2z_bench_amazon500_notiles.err
python - <<'PY'
import json
for c in ['enron50','amazon500','amazon500_notiles']:
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2z_bench_{c}.json') if l.startswith('{')][-1])
        print(c, 'ms/step %.4f kernel %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']), d['roofline']['tiles'])
    except Exception as e:
        print(c, 'FAILED', e)
PY
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-init-a --no-traffic > gpurun_out/r2z_launches_bench.log 2>&1; grep -c tile_step_kernel gpurun_out/r2z_launches.csv
echo "== ncu --set full of the step kernel"; BIGCLAM_AB_SPARSE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_step_kernel --launch-skip 6 --launch-count 1 -f -o gpurun_out/r2z_prof_tile python tools/profile_step.py 200 8 2 > gpurun_out/r2z_ncu.log 2>&1; tail -2 gpurun_out/r2z_ncu.log
