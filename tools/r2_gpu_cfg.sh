#!/bin/bash
# 1-GPU bench lines of the other BASELINE configs (+ a reduced R-MAT as a smoke test of the big one)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== amazon200"; timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2c_amazon200_n1.json 2> gpurun_out/r2c_amazon200_n1.err; tail -c 400 gpurun_out/r2c_amazon200_n1.json; tail -2 gpurun_out/r2c_amazon200_n1.err
echo "== amazon500"; timeout 600 python bench.py --config amazon500 --steps 30 --warmup 5 --no-init-a > gpurun_out/r2c_amazon500_n1.json 2> gpurun_out/r2c_amazon500_n1.err; tail -c 400 gpurun_out/r2c_amazon500_n1.json; tail -2 gpurun_out/r2c_amazon500_n1.err
echo "== enron50"; timeout 600 python bench.py --config enron50 --steps 50 --warmup 5 > gpurun_out/r2c_enron50_n1.json 2> gpurun_out/r2c_enron50_n1.err; tail -c 400 gpurun_out/r2c_enron50_n1.json; tail -2 gpurun_out/r2c_enron50_n1.err
echo "== rmat 1M/10M K=1000 (smoke test of config 5)"; timeout 900 python bench.py --graph rmat:1000000:10000000 --k 1000 --steps 10 --warmup 3 --no-cpu --no-init-a --no-traffic > gpurun_out/r2c_rmat1m_n1.json 2> gpurun_out/r2c_rmat1m_n1.err; tail -c 600 gpurun_out/r2c_rmat1m_n1.json; tail -3 gpurun_out/r2c_rmat1m_n1.err
echo "== GPU conductance tests"; timeout 600 python -m pytest tests/test_gpu_init.py -m gpu -q -s 2>&1 | tail -4
