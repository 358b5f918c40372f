# A/B of the fast-FIR convolutions against the direct one (LSIM_CONV_NO_FF=1): shapes 0-2 row split, 3-4 both directions
mkdir -p gpurun_out
{
LSIM_FF_SHAPE=3 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in noff 3 4; do
  echo "== shape $v"
  unset LSIM_CONV_NO_FF LSIM_FF_SHAPE
  if [ $v = noff ]; then export LSIM_CONV_NO_FF=1; else export LSIM_FF_SHAPE=$v; fi
  python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
export LSIM_FF_SHAPE=3
bash tools/ncu_profile.sh r2w k_conv_tiled_fxy > /dev/null 2>&1
head -60 gpurun_out/k_conv_tiled_fxy_r2w_summary.txt
} 2>&1 | tee gpurun_out/r2_variants8.txt
