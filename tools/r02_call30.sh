#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/c30; mkdir -p $O
timeout -k 5 1200 python -m pytest tests/ -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 $O/pytest_all.log
for wl in C3 C5 C4; do for s in 1 2; do
BLSTM_T2_SUB=$s timeout -k 5 600 python bench.py --workload $wl --steps 8 --warmup 3 > $O/bench_${wl}_sub$s.json 2> $O/bench_${wl}_sub$s.err; echo "$wl sub$s rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c30/bench_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], "value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v["ms_per_step"],3) for k,v in d["kernel_classes"].items()}, {k:v for k,v in d["config"]["plan"].items() if k in ("fwd_G","fwd_C","fwd_CL","fwd_nsub","bwd_G","bwd_C","bwd_nsub","fwd_kernel")})
    except Exception as e:
        print(f, "ERR", e)
PY
