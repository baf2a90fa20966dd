for ctas in default 4 8 16; do
  if [ "$ctas" = "default" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 \
    bench.py --gpus 2 --workload cfg3 --steps 8 --warmup 5 --no-profile > gpurun_out/tune_cfg3_ctas_$ctas.json 2> gpurun_out/tune_cfg3_ctas_$ctas.err
  echo "cfg3 x2 NCCL_MAX_CTAS=$ctas: $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/tune_cfg3_ctas_$ctas.json').read().strip().splitlines()[-1]); print(round(d['value'],1), 'vol/s', round(d['ms_per_step'],2), 'ms', d.get('ddp_mode'))
except Exception as e: print('no line', e)")"
done
unset NCCL_MAX_CTAS
CAVIT_DDP_MODE=post timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --workload cfg3 --steps 8 --warmup 5 --no-profile | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('post mode', round(d['value'],1), round(d['ms_per_step'],2), d.get('ddp_mode'))"
timeout 120 python bench.py --workload cfg3 --steps 8 --warmup 5 --no-cpu-baseline --no-configs0 --no-profile | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('1gpu', round(d['value'],1), round(d['ms_per_step'],2))"
python tools/hbm_bw.py
