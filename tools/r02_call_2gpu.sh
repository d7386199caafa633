#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/g2; mkdir -p $O
run() { # name, env..., -- args
  name=$1; shift
  env "$@" timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 $EXTRA > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"
}
EXTRA="" run c2_overlap4 BLSTM_COMM_MODE=overlap BLSTM_COMM_MAX_CTAS=4
EXTRA="" run c2_grouped BLSTM_COMM_MODE=grouped


EXTRA="--workload C5" run c5_overlap16 BLSTM_COMM_MODE=overlap BLSTM_COMM_MAX_CTAS=16
EXTRA="--workload C5" run c5_grouped BLSTM_COMM_MODE=grouped


python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/g2/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "gpus", d["n_gpus"], "value %.0f e2e %.0f ms/step %.3f dev %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["device_ms_per_step"]), "dp_parity", d.get("dp_parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $O/c2_overlap4.err
