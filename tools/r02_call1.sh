#!/bin/bash
# Round-2 GPU call 1: protocol / MMA / gate-math probes, TF32 peak, the new parity tests on the round-1 kernels, baseline bench lines.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/c1; mkdir -p $O
nvidia-smi > $O/smi.txt 2>&1
timeout -k 5 180 tools/micro/exchange_probe > $O/exchange_probe.txt 2>&1; echo "exchange_probe rc=$?"
timeout -k 5 60 tools/micro/tcgen05_f16_step > $O/tcgen05_f16_step.txt 2>&1; echo "f16 rc=$?"
timeout -k 5 60 tools/micro/tcgen05_ts_step > $O/tcgen05_ts_step.txt 2>&1; echo "ts rc=$?"
timeout -k 5 60 tools/micro/gate_math_probe > $O/gate_math_probe.txt 2>&1; echo "gate rc=$?"
timeout -k 5 120 python tools/measure_tf32_peak.py > $O/tf32_peak.json 2> $O/tf32_peak.err; echo "tf32 rc=$?"
timeout -k 5 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_parity.py -m gpu -q -s -k "reference_layout or long_sequences" > $O/pytest_new.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_new.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout -k 5 600 python bench.py --workload C5 --steps 6 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"
timeout -k 5 300 python bench.py --workload C3 --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 rc=$?"
cat $O/exchange_probe.txt $O/tcgen05_f16_step.txt $O/gate_math_probe.txt
