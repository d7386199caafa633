#!/bin/bash
# Round-2 multi-GPU numbers on one 8-GPU box: C2 / C5 at 8 GPUs, C3 at 2 / 4 / 8 (weak scaling: per-GPU parallel_sequences fixed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/g8; mkdir -p $O
run() { # name, gpus, extra args
  name=$1; n=$2; shift 2
  timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 10 --warmup 3 "$@" 2> $O/$name.err | tail -1 > $O/$name.json; echo "$name rc=$?"
}
run c2_8gpu 8
run c5_8gpu 8 --workload C5
run c3_8gpu 8 --workload C3
run c3_4gpu 4 --workload C3
run c3_2gpu 2 --workload C3
run c5_4gpu 4 --workload C5
BLSTM_COMM_MODE=grouped run c5_8gpu_grouped 8 --workload C5
timeout -k 5 300 python bench.py --workload C3 --steps 10 --warmup 3 2>/dev/null | tail -1 > $O/c3_1gpu.json
timeout -k 5 300 python bench.py --workload C5 --steps 10 --warmup 3 2>/dev/null | tail -1 > $O/c5_1gpu.json
timeout -k 5 300 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > $O/c2_1gpu.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/g8/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "gpus", d["n_gpus"], "value %.0f e2e %.0f ms/step %.3f dev %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["device_ms_per_step"]), "dp_parity", d.get("dp_parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
