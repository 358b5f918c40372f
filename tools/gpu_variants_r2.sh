mkdir -p gpurun_out
python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -x 2>&1 | tail -6
for f in jaxtronomy_b200/csrc/liblsim.so build_variants/*.so; do echo "== $f"; LSIM_LIB=$PWD/$f python bench.py --no-cpu --no-secondary --steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['avg_launch_ms'], d['roofline']['kernel_share_of_step'])"; done 2>&1 | tee gpurun_out/r2_variants3.txt
python tools/bench_configs.py c3 c5 2>&1 | tail -3 | cut -c1-400 | tee -a gpurun_out/r2_variants3.txt
LSIM_NO_FOLD=1 python tools/bench_configs.py c3 2>&1 | tail -2 | cut -c1-300 | tee -a gpurun_out/r2_variants3.txt
