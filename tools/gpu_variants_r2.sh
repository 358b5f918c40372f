mkdir -p gpurun_out
for f in jaxtronomy_b200/csrc/liblsim.so build_variants/*.so jaxtronomy_b200/csrc/liblsim.so; do echo "== $f"; LSIM_LIB=$PWD/$f python bench.py --no-cpu --no-secondary --steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['avg_launch_ms'], d['roofline']['kernel_share_of_step'])"; done 2>&1 | tee gpurun_out/r2_variants4.txt
