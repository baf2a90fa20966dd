#!/bin/bash
# Round profile: bench line, ncu launch list of one training step, ncu --set full captures of the top kernels.
# Run under gpurun from the repo root; everything lands in gpurun_out/ (copy the summaries into profiles/).
TAG=${1:-r01c}
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
python bench.py --steps 10 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
python tools/one_step.py > $O/plain_step.log 2>&1 && \
  ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_$TAG.csv python tools/one_step.py > $O/ncu_step.log 2>&1
python tools/one_gemm.py fc1 > $O/plain_fc1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o $O/prof_gemm_fc1_$TAG -f python tools/one_gemm.py fc1 > $O/ncu_fc1.log 2>&1
python tools/one_attn.py > $O/plain_attn.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 2 -o $O/prof_attn_$TAG -f python tools/one_attn.py > $O/ncu_attn.log 2>&1
python tools/bench_kernels.py --only ln > $O/plain_ln.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:ln_fwd -s 3 -c 1 -o $O/prof_ln_fwd_$TAG -f python tools/bench_kernels.py --only ln > $O/ncu_lnf.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:ln_bwd -s 3 -c 1 -o $O/prof_ln_bwd_$TAG -f python tools/bench_kernels.py --only ln > $O/ncu_lnb.log 2>&1
for wl in cfg1 cfg3 cfg5; do
  python bench.py --workload $wl --steps 5 --warmup 5 --no-cpu-baseline > $O/bench_${wl}_$TAG.json 2> $O/bench_${wl}_$TAG.err
done
python tools/bench_kernels.py > $O/kbench_$TAG.log 2>&1
ls -la $O | tail -20
