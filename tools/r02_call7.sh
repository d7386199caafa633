#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/c9; mkdir -p $O
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -m gpu -q -x -s -k "scalar or families or geometries or tensor_memory or long_sequences or c2_network or c5_lvcsr or c3_chime or reference_layout or c4_chime or extreme" > $O/pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -4 $O/pytest_a.log
timeout -k 5 120 python tools/trace_recurrent.py 250 100 300 > $O/trace_tm2_sub2.txt 2>&1; echo "trace rc=$?"; cat $O/trace_tm2_sub2.txt
BLSTM_T2_SUB=1 timeout -k 5 120 python tools/trace_recurrent.py 250 100 300 > $O/trace_tm2_sub1.txt 2>&1; cat $O/trace_tm2_sub1.txt
for sub in 2 1; do
BLSTM_T2_SUB=$sub timeout -k 5 300 python bench.py --steps 10 --warmup 3 > $O/bench_c2_sub$sub.json 2> $O/bench_c2_sub$sub.err; echo "c2 sub$sub rc=$?"
done
timeout -k 5 600 python bench.py --workload C5 --steps 6 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"
timeout -k 5 300 python bench.py --workload C3 --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 rc=$?"
BLSTM_T2_SUB=1 timeout -k 5 300 python bench.py --workload C3 --steps 10 --warmup 3 > $O/bench_c3_sub1.json 2> $O/bench_c3_sub1.err; echo "c3 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c9/bench_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], "value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v["ms_per_step"],3) for k,v in d["kernel_classes"].items()}, d["config"]["plan"])
    except Exception as e:
        print(f, "ERR", e)
PY
