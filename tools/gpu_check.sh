# GPU round trip used during development: parity tests, then the headline bench (and the opt-in fp32 variant).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --batch 16384 --no-cpu --steps 5 --warmup 3 | tail -1
python bench.py --batch 16384 --no-cpu --steps 5 --warmup 3 --precision f32 | tail -1
