#!/usr/bin/env python
"""bench.py -- training frames/sec of the TIMIT-shape BLSTM (BASELINE.json configs[1]) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3                      (N>1: launched under torch.distributed.run)
    python bench.py --impl reference --gpus 1 --steps 4 --warmup 1       (the reference's own CPU path, rank 0 only)

A "step" is one pass of the hot path over one fraction (S parallel sequences per GPU): loadSequences (H2D) ->
forward -> error -> backward -> per-layer gradient all-reduce (N>1) -> SGD+momentum update.
  value : valid frames/s with the fraction already resident in HBM (CUDA-event time of forward..update, summed over K steps)
  e2e   : valid frames/s through the host C ABI with HOST (pinned) buffers: H2D of the fraction and D2H of the objective
          inside the timed region; K steps bracketed by barrier + synchronize, max over ranks
  roofline / kernel_classes : per-kernel-class device time measured live with CUDA events on the launch stream
  cpu_baseline : the reference's CPU path (oracle/_ref, built from /root/reference) on a bounded sample: its OpenMP-host-backend
          build on all host threads, with the rate of its own single-threaded build configuration beside it
Synthetic data: frames N(0,1), weights U(-0.1,0.1), lengths clip(lognormal(ln 290, .35), 90, 780), lr 1e-4, momentum .9.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "lstm-rnn_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)

import synth  # noqa: E402

METRIC, UNIT = "train_frames_per_sec", "frames/s"
WORKLOAD = {
    "C2": "C2 TIMIT-shape deep BLSTM: input 123 -> 3 x blstm 500 (250 cells/direction) -> softmax 183 -> multiclass CE, "
          "parallel_sequences=100 per GPU, synthetic frames, lengths clip(lognormal(ln290,.35),90,780)",
    # the other BASELINE.json configs are parity-test cases (tests/test_gpu_parity.py); --workload measures them on request
    "C3": "C3 CHiME recognition BLSTM: input 39 -> blstm 156 -> blstm 300 -> blstm 102 -> softmax 51 -> multiclass CE, "
          "parallel_sequences=50 per GPU, synthetic frames, lengths U[113,152]",
    "C4": "C4 CHiME autoencoding BLSTM: input 39 -> blstm 156 -> blstm 256 -> blstm 156 -> feedforward_identity 39 -> sse, "
          "truncate_seq=64, parallel_sequences=50 per GPU, synthetic noisy/clean pairs, lengths U[113,152]",
    "C5": "C5 LVCSR-shape BLSTM: input 123 -> 5 x blstm 1024 (512 cells/direction) -> softmax 8000 -> multiclass CE, truncate_seq=500, "
          "parallel_sequences=16 per GPU (128 on 8 GPUs), synthetic frames, lengths clip(lognormal(ln800,.4),200,2000)",
}
WORKLOAD_SEED = {"C2": 2, "C3": 3, "C4": 4, "C5": 5}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    if os.path.exists(path):
        d = json.load(open(path))
        out = {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
               "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    # dense TF32 peak measured by tools/measure_tf32_peak.py the way the driver measured bf16 (cuBLAS 8192^3, burst / 4 s sustained)
    tpath = os.path.join(ROOT, "profiles", "r02_tf32_peak.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        out.update(tf32_tflops=float(t["tf32_tflops"]), tf32_tflops_sustained=float(t["tf32_tflops_sustained"]))
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 7 and s[3 + i].lower().startswith("active")})
        busy = [x for x in sm if mx and x > 0.5 * mx[0]] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------- workload
def c2_sequences(num_seqs, seed=2):
    cfg = synth.config("C2")
    lengths = synth.sequence_lengths(cfg, num_seqs, seed)
    xs, cs, _ = synth.make_sequences(lengths, 123, seed, classes=183)
    return cfg, lengths, xs, cs


def workload_sequences(name, num_seqs):
    """Synthetic sequences of a BASELINE.json config: (cfg, lengths, inputs, class targets | None, dense targets | None,
    input size, output size).  For C2 this is exactly c2_sequences()."""
    cfg = synth.config(name)
    seed = WORKLOAD_SEED[name]
    layers = json.loads(cfg["net"])["layers"]
    P, O = layers[0]["size"], layers[-1]["size"]
    lengths = synth.sequence_lengths(cfg, num_seqs, seed)
    if cfg["classes"]:
        xs, cs, ts = synth.make_sequences(lengths, P, seed, classes=cfg["classes"])
    else:
        xs, cs, ts = synth.make_sequences(lengths, P, seed, target_size=O)
    return cfg, lengths, xs, cs, ts, P, O


def layer_shapes(net_json):
    layers = json.loads(net_json)["layers"]
    return [(ly["type"], ly["size"], layers[i - 1]["size"] if i else 0) for i, ly in enumerate(layers)]


def algorithmic_work(net_json, slots):
    """Per fraction of `slots` pattern slots: HBM bytes of the recurrent kernels, flops of the GEMMs."""
    fwd_b = bwd_b = flops = 0.0
    first = True
    for t, L, P in layer_shapes(net_json):
        if t in ("blstm", "lstm"):
            ndir = 2 if t == "blstm" else 1
            fwd_b += 40.0 * L * slots            # read 4 pre-acts, write 4 acts + c + h
            bwd_b += 48.0 * L * slots            # read 4 acts + c + c_prev + e, write 4 deltas + eps_c
            H = L // ndir
            flops += slots * (8.0 * L * P + (0 if first else 8.0 * L * P) + 8.0 * L * P + 8.0 * L * H)   # proj, input-err, dW_in, dW_rec
            first = False
        elif t.startswith("feedforward") or t == "softmax":
            flops += slots * 6.0 * L * P
    return fwd_b, bwd_b, flops


# ------------------------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_reference_rate(steps, warmup, S=8, T=40, seed=2, threads=1):
    """The reference's own CPU path (NeuralNetwork<Cpu>, --cuda false) on a bounded sample of the C2 workload.
    threads == 1: the reference as its own build makes it (sequential host Thrust).  threads > 1: the same sources built with
    Thrust's OpenMP host backend (oracle/_ref/libcurrennt_ref_omp.so) -- every host thread the reference can be given."""
    from oracle import pyoracle
    cfg = synth.config("C2")
    net_json = cfg["net"]
    omp = threads > 1 and pyoracle.ref_omp_available()
    if not omp:
        threads = 1
    kind = "reference" if pyoracle.ref_available() else "port"
    rng = np.random.default_rng(seed)
    lengths = np.clip(np.round(rng.lognormal(np.log(0.8 * T), 0.25, S)), 4, T).astype(int)
    lengths[-1] = T
    xs, cs, _ = synth.make_sequences(lengths, 123, seed, classes=183)
    frac = pyoracle.make_fraction(xs, S, 0, seq_classes=cs, O=183)
    if omp:
        pyoracle.set_omp_threads(threads)
        net = pyoracle.RefNet(net_json, S, T, omp=True)
    else:
        net = (pyoracle.RefNet if kind == "reference" else pyoracle.OracleNet)(net_json, S, T)
    for i, w in enumerate(synth.init_weights(net_json, seed + 1)):
        if len(w):
            net.set_weights(i, w)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        net.load_fraction(frac)
        net.forward()
        net.calculate_error()
        net.count_correct()
        net.backward()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    sample = ("C2 network, one fraction S=%d T=%d (%d valid frames of %d slots), %d passes of loadSequences+forward+error+backward "
              "(SGD update excluded: 3.85M weights, <1%% of the step); %s" %
              (S, T, frac.valid_frames, frac.N, steps,
               "Thrust OpenMP host backend build of the unmodified sources, %d threads" % threads if omp
               else "the reference's own build configuration: sequential host Thrust, 1 thread"))
    return {"value": frac.valid_frames * steps / total, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
            "ms_per_step": 1e3 * total / steps, "frames_per_step": frac.valid_frames}


def cpu_reference_best(steps, warmup):
    """The reference with all the host threads it can use: the OpenMP build at the best of a few thread counts (one probe pass
    each), next to the single-threaded configuration its own CMakeLists produces.  Returns (best, single_thread)."""
    from oracle import pyoracle
    single = cpu_reference_rate(min(steps, 4), 1)
    if not pyoracle.ref_omp_available():
        return single, single
    ncpu = os.cpu_count() or 1
    cands = sorted({n for n in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 < n <= ncpu})
    if not cands:
        return single, single
    try:
        probe = {n: cpu_reference_rate(1, 1, threads=n)["value"] for n in cands}
        best_n = max(probe, key=probe.get)
        best = cpu_reference_rate(steps, warmup, threads=best_n)
    except OSError as e:                               # no OpenMP runtime on this host: the single-threaded build is the baseline
        print("bench: OpenMP reference build unusable (%s)" % e, file=sys.stderr)
        return single, single
    return (best if best["value"] > single["value"] else single), single


def run_reference(args, rank):
    if rank != 0:
        return
    steps = args.steps
    base, single = cpu_reference_best(steps, max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD["C2"], "reference_sample": base["sample"]},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "single_thread": {k: single[k] for k in ("value", "unit", "cores", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import currennt_b200 as cb

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    ctx = cb.Context(local_rank, stream.cuda_stream)
    if args.mode == "fast":
        ctx.set_gemm_mode(1)
    k, h = cb.libs()

    K, W = args.steps, args.warmup
    wl = args.workload
    S = 16 if wl == "C5" else synth.config(wl)["S"]              # per GPU (weak scaling); C5's 128 is the 8-GPU global figure
    cfg, lengths, xs, cs, ts, P_in, O_out = workload_sequences(wl, (K + W) * S * world)
    classification = cs is not None
    net_json = cfg["net"]
    ds = cb.DataSet(ctx, xs, S, seq_classes=cs, seq_targets=ts, O=O_out, truncate=cfg["truncate"], training=True, rank=rank, world=world)
    del xs, cs, ts
    net = cb.Net(ctx, net_json, S, ds.max_len)
    for i, w in enumerate(synth.init_weights(net_json, 3)):
        if len(w):
            net.set_weights(i, w)
    comm = None
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (ctypes.c_ubyte * 128)()
            assert k.bl_comm_unique_id(raw) == 0, k.bl_last_error(None)
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        raw = (ctypes.c_ubyte * 128)(*idbuf.cpu().tolist())
        cp = ctypes.c_void_p()
        ctx.check(k.bl_comm_create(ctx.p, rank, world, raw, ctypes.byref(cp)))
        comm = cp
        net.set_comm(comm)
    opt = cb.Optimizer(net, 1e-4, 0.9, hybrid=True)

    fracs = []
    while True:
        f = ds.next_fraction()
        if f is None:
            break
        fracs.append(f)
    order = np.random.default_rng(5).permutation(len(fracs))       # shuffle_fractions=true (examples/phoneme_recognition_timit/config.cfg)
    fracs = [fracs[i] for i in order]
    warm, timed = fracs[:W], fracs[W:W + K]
    warm = [max(timed, key=lambda f: f.N)] + warm          # sizes every scratch buffer before the timed passes
    assert len(timed) == K, "not enough fractions"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for f in warm:
            opt.train_fraction(f)
        barrier()

        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()

        # ---- pass 1 (`value`): fraction already resident in HBM when the timed region of a step starts: the H2D copies of
        # loadSequences are enqueued first, then CUDA events bracket forward .. weight update on the launch stream
        barrier()
        dev_ms = 0.0
        evs = []
        for f in timed:
            net.load_fraction(f)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            net.forward()
            net.calculate_error()
            if classification:
                net.count_correct()
            net.backward()
            opt.update_weights()
            e1.record(stream)
            evs.append((e0, e1))
        torch.cuda.synchronize()
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)

        # ---- pass 2: end to end through the host ABI with pinned host buffers, K steps in one timed region
        barrier()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        frames = 0
        for f in timed:
            _, _, n = opt.train_fraction(f)
            frames += n
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        launches = ctx.launches - l0
        clocks = sampler.summary() if rank == 0 else None

        # ---- pass 3 (not part of any reported throughput): per-kernel-class device time from the library's own event
        # brackets around its launches, for the roofline line
        k.bl_ctx_timing_enable(ctx.p, 1)
        ms = (ctypes.c_double * 4)()
        cnt = (ctypes.c_long * 4)()
        k.bl_ctx_timing_read(ctx.p, ms, cnt)                       # reset
        for f in timed:
            opt.train_fraction(f)
        k.bl_ctx_timing_read(ctx.p, ms, cnt)
        class_ms, class_cnt = [float(x) for x in ms], [int(x) for x in cnt]
        k.bl_ctx_timing_enable(ctx.p, 0)

    # data-parallel parity: after the same K steps every replica must hold bit-identical weights
    dp_parity = None
    if dist is not None:
        import zlib
        crc = [zlib.crc32(net.get_weights(i).tobytes()) for i in range(1, len(layer_shapes(net_json)) - 1)]
        mine = torch.tensor(crc, dtype=torch.int64, device="cuda")
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        dp_parity = all(bool(torch.equal(allc[0], c)) for c in allc)
        assert dp_parity, "data-parallel replicas diverged: weight checksums differ across ranks"

    tot = torch.tensor([float(frames), dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_frames, dev_ms, e2e_ms = float(tot[0]), float(mx[1]), float(mx[2])
    else:
        total_frames = float(frames)

    if rank == 0:
        pk = peaks()
        slots = sum(f.N for f in timed)
        fwd_b, bwd_b, flops = algorithmic_work(net_json, slots)
        names = ["gemm", "lstm_recurrent_fwd", "lstm_bptt", "elementwise"]
        share = [m / max(sum(class_ms), 1e-9) for m in class_ms]
        classes = {}
        for i, nm in enumerate(names):
            d = {"ms_per_step": class_ms[i] / K, "launch_groups_per_step": class_cnt[i] / K, "share_of_kernel_time": share[i]}
            if nm == "gemm":
                tf = flops / (class_ms[i] * 1e-3) / 1e12 if class_ms[i] else None
                d.update(achieved_tflops=tf, frac_of_bf16_tensor_peak=tf / pk["bf16_tflops_sustained"] if tf else None)
                if tf and "tf32_tflops_sustained" in pk:
                    # algorithmic flops against the measured dense TF32 peak; the strict mode issues 3 TF32 MMAs per useful one
                    d.update(frac_of_tf32_tensor_peak=tf / pk["tf32_tflops_sustained"], tf32_issue_frac=(1.0 if args.mode == "fast" else 3.0) * tf / pk["tf32_tflops_sustained"],
                             tf32_peak_tflops_sustained=pk["tf32_tflops_sustained"])
                    # the GEMMs alternate with the low-power recurrent kernels, so they run nearer the burst (unthrottled) peak than the
                    # 4-s sustained one: against the sustained figure the issue fraction can exceed 1 (C5), so the burst one is given too
                    d.update(tf32_peak_tflops_burst=pk["tf32_tflops"], tf32_issue_frac_of_burst=d["tf32_issue_frac"] * pk["tf32_tflops_sustained"] / pk["tf32_tflops"])
            if nm == "lstm_recurrent_fwd":
                d.update(achieved_gbs=fwd_b / (class_ms[i] * 1e-3) / 1e9 if class_ms[i] else None)
            if nm == "lstm_bptt":
                d.update(achieved_gbs=bwd_b / (class_ms[i] * 1e-3) / 1e9 if class_ms[i] else None)
            classes[nm] = d
        # dominant kernel = the persistent recurrent kernels (forward + BPTT), bounded by HBM per SURVEY.md 8d
        rec_ms = class_ms[1] + class_ms[2]
        n_launch = class_cnt[1] + class_cnt[2]
        achieved = (fwd_b + bwd_b) / (rec_ms * 1e-3) / 1e9
        traffic = traffic_note = None
        tpath = os.path.join(ROOT, "profiles", "recurrent_traffic.json")
        if os.path.exists(tpath):        # DRAM bytes per slot and unit of layer size from the committed ncu --set full captures
            tj = json.load(open(tpath))
            traffic_note = "scaled from the ncu --set full captures of " + tj.get("captured_kernels", "the kernels named in `kernel`") + " (profiles/recurrent_traffic.json)"
            traffic = (fwd_b / 40.0 * tj["fwd_dram_bytes_per_slot_per_layer_unit"] + bwd_b / 48.0 * tj["bwd_dram_bytes_per_slot_per_layer_unit"]) / max(n_launch, 1)
        plan = net.plan_info(2)
        kfam = {"tm2": "tm2", "tmem": "tmem", "registers": "reg", "smem": "persistent"}
        roofline = {"kernel": "lstm_fwd_%s_kernel + lstm_bwd_%s_kernel" % (kfam[plan["fwd_kernel"]], kfam[plan["bwd_kernel"]]), "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / pk["hbm_gbs"], "peak_source": pk["source"], "traffic": traffic, "traffic_source": traffic_note,
                    "algorithmic_bytes_per_launch": (fwd_b + bwd_b) / max(n_launch, 1), "avg_launch_ms": rec_ms / max(n_launch, 1),
                    "share_of_kernel_time": share[1] + share[2]}
        # inputs + targets (int class or dense row) + one patTypes byte per slot and layer
        h2d = sum(f.N * (P_in * 4 + (4 if classification else O_out * 4)) for f in timed) / K + sum(f.N for f in timed) / K * len(layer_shapes(net_json))
        base = single = None
        if world == 1 and wl == "C2":
            # in its own process: an OpenMP runtime that shares a process with torch's runs the reference build several times slower
            try:
                sub = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "4", "--warmup", "1"],
                                     capture_output=True, text=True, timeout=900)
                ref = json.loads(sub.stdout.strip().splitlines()[-1])
                base, single = ref["cpu_baseline"], ref["single_thread"]
            except Exception as e:
                print("bench: reference subprocess failed (%s); timing it in-process" % e, file=sys.stderr)
                base, single = cpu_reference_best(4, 1)
        line = {"metric": METRIC, "value": total_frames / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": e2e_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD[wl], "parallel_sequences_per_gpu": S, "gemm_mode": args.mode,
                           "frames_per_step": total_frames / K, "slots_per_step_per_gpu": slots / K,
                           "l2": "inputs larger than L2: every step touches ~%.1f GB of activations/deltas per GPU"
                                 % (slots / K * sum(4 * L for t, L, _ in layer_shapes(net_json) if t in ("lstm", "blstm")) * 4 * 2 / 1e9),
                           "plan": plan},
                "e2e": {"value": total_frames / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8},
                "device_ms_per_step": dev_ms / K, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
                "kernel_classes": classes,
                "notes": "value = fraction already resident in HBM when the timed region of a step starts (the bench contract); e2e = the "
                         "same K steps through the host C ABI with pinned host buffers, H2D of the fraction and D2H of the objective inside "
                         "the timed region -- the headline"}
        if dp_parity is not None:
            line["dp_parity"] = dp_parity
            nw = sum(synth.layer_num_weights(t, L, P) for t, L, P in layer_shapes(net_json))
            mode = os.environ.get("BLSTM_COMM_MODE", "overlap")
            line["collective"] = {"kind": "ncclAllReduce(sum, fp32) of every layer's weightUpdates, in place", "bytes_per_step": 4 * nw,
                                  "schedule": ("per layer on a side stream as soon as that layer's backward pass is enqueued, joined before the "
                                               "weight update; communicator capped to %s CTAs (the SMs the persistent kernels leave free)"
                                               % os.environ.get("BLSTM_COMM_MAX_CTAS", "4")) if mode != "grouped"
                                  else "one grouped call after the whole backward pass"}
        if base:
            line["cpu_baseline"] = {kk: base[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["single_thread_value"] = single["value"]
        print(json.dumps(line), flush=True)

    if comm is not None:
        k.bl_comm_destroy(comm)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOAD),
                    help="BASELINE.json config to measure; the default C2 is the one the metric is quoted on (the reference arm and cpu_baseline are C2 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if args.workload != "C2":
            raise SystemExit("bench.py --impl reference times the C2 workload only")
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
