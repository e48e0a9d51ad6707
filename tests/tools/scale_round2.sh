# End-of-round multi-GPU lines: bash tests/tools/scale_round2.sh N  (inside `gpurun --gpus N`)
N=$1
mkdir -p gpurun_out
run() {  # name, extra args
  name=$1; shift
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 "$@" > gpurun_out/${name}_1gpu.json 2> gpurun_out/${name}_1gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/${name}_${N}gpu.json 2> gpurun_out/${name}_${N}gpu.err
  fi
  echo "$name N=$N rc=$?"; tail -c 300 gpurun_out/${name}_${N}gpu.json | head -c 300; echo
}
run r02z_cfg5 --config cfg5 --steps 2 --warmup 2 --no-cpu
if [ "$N" != 1 ] && [ -z "$SKIP_CFG4" ]; then run r02z_cfg4 --steps 3 --warmup 3 --no-cpu; fi
