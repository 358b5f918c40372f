# chunk size (parameter sets per launch = SMs x LSIM_CHUNK_MULT) against the headline
mkdir -p gpurun_out
{
for v in 8 4 16 32 64; do
  echo "== LSIM_CHUNK_MULT=$v"
  LSIM_CHUNK_MULT=$v python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_share_of_step'])"
done
} 2>&1 | tee gpurun_out/r2_variants9.txt
