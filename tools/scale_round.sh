#!/bin/bash
# Data-parallel scaling runs of the BASELINE.json configs that name multi-GPU (cfg3: DDP on 8 B200; cfg5: 1/2/4/8 B200) on ONE
# 8-GPU box, most important runs first; every run under its own timeout. Usage (gpurun --gpus 8): bash tools/scale_round.sh r02
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
run() {  # workload gpus extra-args...
  wl=$1; n=$2; shift 2
  f=$out/scale_${tag}_${wl}_${n}gpu.json
  if [ "$n" = "1" ]; then
    timeout 170 python bench.py --workload $wl --steps 8 --warmup 5 --no-cpu-baseline --no-configs0 "$@" > $f 2> $out/scale_${tag}_${wl}_${n}gpu.err
  else
    timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --workload $wl --steps 8 --warmup 5 --no-profile "$@" > $f 2> $out/scale_${tag}_${wl}_${n}gpu.err
  fi
  echo "$wl x$n rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1), 'vol/s', round(d['ms_per_step'],2), 'ms', d.get('ddp_mode'), 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('no line', e)")"
}
NCCL_DEBUG=INFO run cfg3 8
run cfg3 1
run cfg5 8
run cfg5 1
run cfg3 4
run cfg3 2
run cfg5 4
run cfg5 2
grep -h -m1 "NVLS\|nvls" $out/nccl_*.log | head -3
