# ncu captures of the three kernels of the headline step (one launch each, full set) + the launch list.
# Run on the GPU box: gpurun -- 'bash tools/ncu_profile.sh <tag>'; text summaries land in gpurun_out/ (the
# .ncu-rep files are ~45 MB each and stay on the box).
TAG=${1:-dev}
KERNELS=${2:-"k_image k_conv_tiled_u k_likelihood"}
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 1 --warmup 3 --batch 9472 --no-cpu --no-secondary"
for k in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 -c 1 -f \
      -o /tmp/ncu/${k}_${TAG} $CMD > gpurun_out/ncu_${k}_${TAG}.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/${k}_${TAG}.ncu-rep 30 > gpurun_out/${k}_${TAG}_summary.txt 2>&1
  ncu -i /tmp/ncu/${k}_${TAG}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${k}_${TAG}_source.csv.gz
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD \
    > gpurun_out/ncu_launches_${TAG}.log 2>&1
ls -la gpurun_out
