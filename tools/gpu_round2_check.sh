# Round-2 GPU check: parity tests, bench line, ncu capture of the headline kernels.  gpurun -- 'bash tools/gpu_round2_check.sh <tag> [ncu kernels]'
TAG=${1:-r2}
KERNELS=${2:-"k_image"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -8 gpurun_out/${TAG}_pytest.log
python bench.py --no-secondary > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
if [ -n "$KERNELS" ]; then bash tools/ncu_profile.sh ${TAG} "$KERNELS" > gpurun_out/${TAG}_ncu.log 2>&1; cat gpurun_out/k_image_${TAG}_summary.txt | head -40; fi
