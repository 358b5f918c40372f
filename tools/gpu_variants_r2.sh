# A/B on one box: the shipped library against variant builds (build_variants/*.so, selected with LSIM_LIB)
mkdir -p gpurun_out
{
for f in jaxtronomy_b200/csrc/liblsim.so build_variants/*.so jaxtronomy_b200/csrc/liblsim.so; do
  echo "== $f"
  LSIM_LIB=$PWD/$f python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_ms'])"
done
for f in build_variants/*.so; do LSIM_LIB=$PWD/$f python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -x 2>&1 | tail -1; done
} 2>&1 | tee gpurun_out/r2_variants10.txt
