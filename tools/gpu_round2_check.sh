# Round-2 GPU check: parity tests, smoke, bench line (with secondary configs), ncu captures of the headline kernels.
#   gpurun -- 'bash tools/gpu_round2_check.sh <tag> [ncu kernels]'
TAG=${1:-r2}
KERNELS=${2:-"k_image k_conv_tiled k_likelihood"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); r=d['roofline']; print('value',d['value'],'e2e',d['e2e']['value'],'image ms',r['avg_launch_ms'],'frac',r['frac'],r['kernel_share_of_step']); print(d['cpu_baseline']); print({k:(v.get('ms') or v.get('wall_s') or v) for k,v in d['secondary'].items()})"
if [ -n "$KERNELS" ]; then bash tools/ncu_profile.sh ${TAG} "$KERNELS" > gpurun_out/${TAG}_ncu.log 2>&1; for k in $KERNELS; do head -22 gpurun_out/${k}_${TAG}_summary.txt; done; fi
