#!/bin/bash
# One GPU call that regenerates the raw material of profiles/ (run under gpurun from the repo root):
#   gpurun --timeout 1800 -- 'bash tools/profile_round.sh r01'
# then, back in the container:  python tools/profile_collect.py r01
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
# 1. the bench line itself (never taken under a profiler)
timeout 600 python bench.py 2>$OUT/${TAG}_bench.err | tail -1 > $OUT/${TAG}_bench_1gpu.json
# 2. launch list of the same command (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_launches.log 2>&1
# 3. one full capture of each top kernel
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_ -s 4 -c 1 -f -o $OUT/${TAG}_fwd_full \
    python bench.py --steps 1 --warmup 3 > $OUT/${TAG}_ncu_fwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm_bwd_ -s 4 -c 1 -f -o $OUT/${TAG}_bwd_full \
    python bench.py --steps 1 --warmup 3 > $OUT/${TAG}_ncu_bwd.log 2>&1
# (PROFILE_GEMM=0 skips the GEMM capture when those kernels have not changed since the committed summary)
if [ "${PROFILE_GEMM:-1}" = "1" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:2cta -s 40 -c 13 -f -o $OUT/${TAG}_gemm_full \
    python bench.py --steps 1 --warmup 3 > $OUT/${TAG}_ncu_gemm.log 2>&1
fi
# 4. in-kernel phase trace of the recurrent kernels (one C2 layer, T=300; one C5 layer, T=200)
BLSTM_REC_TRACE=1 timeout 300 python tools/trace_recurrent.py 250 100 300 > $OUT/${TAG}_recurrent_trace.txt 2>&1
BLSTM_REC_TRACE=1 timeout 300 python tools/trace_recurrent.py 512 16 200 > $OUT/${TAG}_recurrent_trace_h512.txt 2>&1
# 5. the micro-benchmarks the kernel design rests on
for p in exchange_probe tcgen05_f16_step gate_math_probe occupancy_probe; do
  echo "==== tools/micro/$p" >> $OUT/${TAG}_probes.txt; timeout 200 tools/micro/$p >> $OUT/${TAG}_probes.txt 2>&1
done
# 6. the other BASELINE.json configs on one GPU
for wl in C3 C4 C5; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 2>/dev/null | tail -1 > $OUT/${TAG}_bench_${wl}_1gpu.json; done
ls -la $OUT | tail -20
