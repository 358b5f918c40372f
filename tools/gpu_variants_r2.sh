# A/B of the specialised k_ray_shoot (LSIM_NO_SPECIALISE=1 runs the interpreted instantiation) on C2 / C4 + parity
mkdir -p gpurun_out
{
for v in 0 1; do
  echo "== LSIM_NO_SPECIALISE=$v"
  if [ $v = 1 ]; then export LSIM_NO_SPECIALISE=1; else unset LSIM_NO_SPECIALISE; fi
  python tools/bench_configs.py c2 c4 2>&1 | grep config | cut -c1-200
done
unset LSIM_NO_SPECIALISE
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_variants6.txt
