#!/bin/bash
# Round-2 profile: ncu launch list of ONE cfg2 training step (eager launches), `ncu --set full` captures of the top / new
# kernels, kernel micro-benchmarks, and a same-box A/B of the fused vs unfused patch embedding. Each ncu run only after the
# same command exited 0 without ncu. Run under gpurun from the repo root; results land in gpurun_out/.
TAG=${1:-r02}
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 120 python tools/one_step.py > $O/plain_step.log 2>&1 && \
  timeout 400 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_$TAG.csv python tools/one_step.py > $O/ncu_step.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:embed_ -c 3 -o $O/prof_embed_$TAG -f python tools/one_step.py > $O/ncu_embed.log 2>&1
timeout 100 python tools/one_gemm.py fc1 > $O/plain_fc1.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o $O/prof_gemm_fc1_$TAG -f python tools/one_gemm.py fc1 > $O/ncu_fc1.log 2>&1
timeout 100 python tools/one_attn.py > $O/plain_attn.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 2 -o $O/prof_attn_$TAG -f python tools/one_attn.py > $O/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:xfold_ -c 2 -o $O/prof_xfold_$TAG -f python tools/one_step.py > $O/ncu_xfold.log 2>&1
timeout 200 python tools/bench_kernels.py > $O/kbench_$TAG.log 2>&1
for f in 1 0; do
  CAVIT_EMBED_FUSED=$f timeout 200 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-configs0 --no-fp32 > $O/bench_embedfused${f}_$TAG.json 2> $O/bench_embedfused${f}_$TAG.err
done
ls -la $O | grep $TAG
