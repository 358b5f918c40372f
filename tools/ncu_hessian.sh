# ncu --set full captures of k_hessian (EPL+SHEAR, 1,024 sets x 32,400 points) and k_source_position (65,536 sets x 4
# images), one launch each after the warm-up launches of tools/bench_hessian.py; summaries into gpurun_out/.
# Not run in round 1 (the GPU budget ended with the event timings of profiles/r1o_hessian_position.md).
TAG=${1:-dev}
mkdir -p gpurun_out /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:k_hessian --launch-skip 3 -c 1 -f -o /tmp/ncu/hessian_${TAG} \
    python tools/bench_hessian.py > gpurun_out/ncu_hessian_${TAG}.log 2>&1
python tools/ncu_summary.py /tmp/ncu/hessian_${TAG}.ncu-rep 30 > gpurun_out/hessian_${TAG}_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_source_position --launch-skip 3 -c 1 -f \
    -o /tmp/ncu/srcpos_${TAG} python tools/bench_hessian.py > gpurun_out/ncu_srcpos_${TAG}.log 2>&1
python tools/ncu_summary.py /tmp/ncu/srcpos_${TAG}.ncu-rep 30 > gpurun_out/source_position_${TAG}_summary.txt 2>&1
