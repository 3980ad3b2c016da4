#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full GPU test suite"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2f_pytest_gpu.log
echo "== enron50"; timeout 600 python bench.py --config enron50 --steps 50 --warmup 5 --no-cpu > gpurun_out/r2f_enron50_n1.json 2> gpurun_out/r2f_enron50_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2f_enron50_n1.json') if l.startswith('{')][-1])
print('enron50 ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], d['roofline']['tiles'], 'initA', d['reference_init_workload'] and d['reference_init_workload']['ms_per_step'])
PY
echo "== amazon200 quick"; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-init-a --no-traffic > gpurun_out/r2f_amazon200_n1.json 2> gpurun_out/r2f_amazon200_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2f_amazon200_n1.json') if l.startswith('{')][-1])
print('amazon200 ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], d['roofline']['tiles'])
PY
echo "== sanitizer"; tools/r2_gpu_sanitize.sh
