#!/bin/bash
# Round-2 GPU call 3: first run of the tm2 kernels -- focused parity tests first (each under timeout), then traces and bench lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/c5; mkdir -p $O
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "scalar or families or geometries or tensor_memory" > $O/pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -4 $O/pytest_a.log
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "long_sequences or c2_network or c5_lvcsr or c3_chime" > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -15 $O/pytest_all.log
timeout -k 5 120 python tools/trace_recurrent.py 250 100 300 > $O/trace_tm2.txt 2>&1; echo "trace rc=$?"; cat $O/trace_tm2.txt
BLSTM_REC_V=3 timeout -k 5 120 python tools/trace_recurrent.py 250 100 300 > $O/trace_tmem.txt 2>&1; cat $O/trace_tmem.txt
timeout -k 5 120 python tools/trace_recurrent.py 512 16 200 > $O/trace_tm2_h512.txt 2>&1; cat $O/trace_tm2_h512.txt
timeout -k 5 300 python bench.py --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout -k 5 600 python bench.py --workload C5 --steps 6 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"
timeout -k 5 300 python bench.py --workload C3 --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
for n in ("c2","c5","c3"):
    try:
        d=json.load(open('gpurun_out/c5/bench_%s.json'%n))
        print(n, "value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v["ms_per_step"],3) for k,v in d["kernel_classes"].items()}, d["config"]["plan"]["fwd_kernel"])
    except Exception as e:
        print(n, "ERR", e)
PY
