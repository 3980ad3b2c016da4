#!/bin/bash
# tools/mg.sh N [exchange] [steps] [layout]: run the N-GPU bench and print the key numbers (layout: dense | sparse)
N=$1; EX=${2:-p2p}
BIGCLAM_EXCHANGE=$EX python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${3:-50} --warmup 5 --layout ${4:-dense} > /tmp/mg_$N.log 2>&1
python - <<PY
import json
line=[l for l in open('/tmp/mg_$N.log') if l.startswith('{"metric')]
if not line:
    print(open('/tmp/mg_$N.log').read()[-1500:])
else:
    d=json.loads(line[-1]); print('N=%d $EX: %.3f ms/step, %.3g edges/s, rank kernel ms %s, rank0 roofline frac %.3f' % (d['n_gpus'], d['ms_per_step'], d['value'], d.get('rank_step_kernel_ms'), d['roofline']['frac']))
PY
