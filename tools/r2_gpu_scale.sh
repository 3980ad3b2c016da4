#!/bin/bash
# tools/r2_gpu_scale.sh N [rmat]: N-GPU lines of the BASELINE configs (run with gpurun --gpus N); logs gpurun_out/r2g_*
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
run() {   # name, extra bench args
  local name=$1; shift
  if [ "$N" = "1" ]; then
    timeout ${TMO:-300} python bench.py --no-cpu --no-init-a --no-traffic "$@" > gpurun_out/r2g_${name}_n$N.json 2> gpurun_out/r2g_${name}_n$N.err
  else
    timeout ${TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r2g_${name}_n$N.json 2> gpurun_out/r2g_${name}_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2g_${name}_n$N.json') if l.startswith('{')][-1])
    print('${name} N=%d: %.4f ms/step  %.4g edges/s  e2e %.4f ms  rank kernels %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('rank_step_kernel_ms', d['roofline'].get('kernel_ms'))))
except Exception as e:
    print('${name} N=$N FAILED', e); print(open('gpurun_out/r2g_${name}_n$N.err').read()[-1500:])
PY
}
nvidia-smi -L | wc -l
if [ "${2:-}" = "rmatonly" ]; then
  TMO=540 run rmat --config rmat --steps 5 --warmup 3
  exit 0
fi
if [ "${2:-}" = "final8" ]; then
  TMO=540 run rmat --config rmat --steps 5 --warmup 3
  TMO=240 run amazon500 --config amazon500 --steps 30 --warmup 5
  TMO=240 run amazon200 --steps 50 --warmup 5
  exit 0
fi
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "$N" != "1" ]; then
  echo "== bigclam_multi_* tests"; timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2g_pytest_multi_n$N.log 2>&1; tail -2 gpurun_out/r2g_pytest_multi_n$N.log
fi
TMO=240 run amazon200 --steps 50 --warmup 5
if [ "${2:-}" = "rmat" ]; then
  TMO=540 run rmat --config rmat --steps 5 --warmup 3
else
  TMO=240 run amazon500 --config amazon500 --steps 30 --warmup 5
fi
