# 1/2/4/8-GPU strong-scaling run of the headline bench on one box (the driver launches it the same way).
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --no-cpu --steps 5 --warmup 3 | tail -1 > gpurun_out/scale_$n.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu --steps 5 --warmup 3 2>gpurun_out/scale_$n.err | tail -1 > gpurun_out/scale_$n.json; fi
  python -c "import json,sys; d=json.loads(open('gpurun_out/scale_$n.json').read()); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
