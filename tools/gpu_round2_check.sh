# Round-2 GPU check: parity tests, bench line, ncu capture of the headline kernels.  gpurun -- 'bash tools/gpu_round2_check.sh <tag> [ncu kernels]'
TAG=${1:-r2}
KERNELS=${2:-"k_image"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -8 gpurun_out/${TAG}_pytest.log
python bench.py --no-secondary > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); r=d['roofline']; print('value',d['value'],'e2e',d['e2e']['value'],'image ms',r['avg_launch_ms'],'frac',r['frac'],r['kernel_share_of_step'])"
LSIM_NO_OVERLAP=1 python bench.py --no-cpu --no-secondary --steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no overlap', d['value'], d['ms_per_step'], d['roofline']['kernel_share_of_step'])"
if [ -n "$KERNELS" ]; then bash tools/ncu_profile.sh ${TAG} "$KERNELS" > gpurun_out/${TAG}_ncu.log 2>&1; for k in $KERNELS; do head -30 gpurun_out/${k}_${TAG}_summary.txt; done; fi
