"""Turns the raw output of tools/profile_round.sh (gpurun_out/) into the committed summaries under profiles/.
Usage: python tools/profile_collect.py r01"""
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")


def run(args):
    return subprocess.run([sys.executable] + args, capture_output=True, text=True, check=True, cwd=root).stdout


bench = json.load(open(os.path.join(src, tag + "_bench_1gpu.json")))
json.dump(bench, open(os.path.join(dst, tag + "_bench_1gpu.json"), "w"), indent=1)
shutil.copy(os.path.join(src, tag + "_launches.csv"), os.path.join(dst, tag + "_launches.csv"))
steps = 10          # bench.py --steps 2 --warmup 3: 4 warm-up steps (the largest fraction first) + 2 steps in each of its value, e2e and timing passes
head = ("# ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 python bench.py --steps 2 --warmup 3\n"
        "# B200, C2 workload, strict mode. Cold-cache serialised times: compare SHARES with bench.py's kernel_classes, not absolutes.\n"
        "# raw csv: profiles/%s_launches.csv; per-step columns divide by the %d steps captured\n" % (tag, steps))
open(os.path.join(dst, tag + "_launch_list_summary.txt"), "w").write(head + run(["tools/launch_summary.py", os.path.join(src, tag + "_launches.csv"), str(steps)]))
fam = {"tm2": "tm2", "tmem": "tmem", "registers": "reg", "smem": "persistent"}
kf, kb = fam[bench["config"]["plan"]["fwd_kernel"]], fam[bench["config"]["plan"]["bwd_kernel"]]
dram = {}
for name, what in (("fwd", "lstm_fwd_%s_kernel (second BLSTM layer of C2, the T=780 fraction: 78 000 slots)" % kf),
                   ("bwd", "lstm_bwd_%s_kernel (same layer and fraction)" % kb),
                   ("gemm", "gemm_tf32_tcgen05_2cta_kernel: the GEMM launches of one training step")):
    rep = os.path.join(src, "%s_%s_full.ncu-rep" % (tag, name))
    if os.path.exists(rep):
        text = run(["tools/ncu_summary.py", rep])
        open(os.path.join(dst, "%s_%s_ncu.txt" % (tag, name)), "w").write(
            "# ncu --set full --clock-control none --import-source on: %s\n# summary by tools/ncu_summary.py (the .ncu-rep stays in gpurun_out/)\n" % what + text)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for line in text.splitlines():
            f = line.split()
            if len(f) >= 4 and f[0] == "DRAM" and f[1] in ("read", "write"):
                tot += float(f[2].replace(",", "")) * scale[f[3]]
        dram[name] = tot
if "fwd" in dram and "bwd" in dram:
    # `ncu -k regex:lstm_fwd_ -s 4 -c 1 python bench.py --steps 1 --warmup 3` captures the 5th launch: second warm-up step (the
    # T=780 fraction), second BLSTM layer (L=500) -> 78 000 slots; bench.py scales these per-slot figures to its own launches
    slots, L = 780 * 100, 500
    json.dump({"source": "profiles/%s_fwd_ncu.txt and profiles/%s_bwd_ncu.txt (ncu --set full, lstm_fwd_%s_kernel / lstm_bwd_%s_kernel): "
                         "dram__bytes_read.sum + dram__bytes_write.sum of one launch, second BLSTM layer of C2 (L=500, S=100) on the T=780 fraction" % (tag, tag, kf, kb),
               "fwd_dram_bytes_per_slot_per_layer_unit": dram["fwd"] / slots / L, "bwd_dram_bytes_per_slot_per_layer_unit": dram["bwd"] / slots / L,
               "fwd_algorithmic": 40.0, "bwd_algorithmic": 48.0}, open(os.path.join(dst, "recurrent_traffic.json"), "w"), indent=1)
tr = os.path.join(src, tag + "_recurrent_trace.txt")
if os.path.exists(tr):
    open(os.path.join(dst, tag + "_recurrent_trace.txt"), "w").write(
        "# BLSTM_REC_TRACE=1 python tools/trace_recurrent.py 250 100 300 on B200: in-kernel clock64 stamps of lstm_fwd_%s_kernel / lstm_bwd_%s_kernel\n" % (kf, kb) + open(tr).read())
print("profiles/ updated:", sorted(f for f in os.listdir(dst) if f.startswith(tag)))
