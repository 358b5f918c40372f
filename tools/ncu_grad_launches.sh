# launch list (ncu gpu__time_duration.sum) of one gradient call at the headline shape, 1184 sets
TAG=${1:-dev}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/grad_launches_${TAG}.csv \
    python tools/bench_grad.py --batch 1184 --steps 1 --warmup 1 > gpurun_out/ncu_grad_launches_${TAG}.log 2>&1
