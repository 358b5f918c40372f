# ncu --set full capture of k_image_grad (one launch, 296 sets at the headline shape); summary into gpurun_out/
TAG=${1:-dev}
mkdir -p gpurun_out /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:k_image_grad --launch-skip 2 -c 1 -f -o /tmp/ncu/grad_${TAG} \
    python tools/bench_grad.py --batch 296 --steps 1 --warmup 1 > gpurun_out/ncu_grad_${TAG}.log 2>&1
python tools/ncu_summary.py /tmp/ncu/grad_${TAG}.ncu-rep 30 > gpurun_out/image_grad_${TAG}_summary.txt 2>&1
