mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python tools/bench_configs.py c2 c3 c4 c5 2>&1 | tail -5 | cut -c1-330 | tee gpurun_out/r2_configs_n.txt
