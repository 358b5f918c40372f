mkdir -p gpurun_out /tmp/p
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/p/dmma_probe tools/dmma_probe.cu && /tmp/p/dmma_probe | tee gpurun_out/r2_dmma_probe.txt
for tol in 1e-14 1e-13 1e-12; do echo "series tol $tol"; LSIM_SERIES_TOL=$tol python bench.py --no-cpu --no-secondary --steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['avg_launch_ms'], d['roofline']['kernel_share_of_step'])"; done 2>&1 | tee gpurun_out/r2_series_tol.txt
LSIM_NO_OVERLAP=1 python bench.py --no-cpu --no-secondary --steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no overlap', d['value'], d['ms_per_step'])" | tee -a gpurun_out/r2_series_tol.txt
