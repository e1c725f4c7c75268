#!/usr/bin/env python
"""bench.py — SSN forward/backward hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fast|exact]
  N>1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): THUMOS14-shape synthetic, per GPU 4 videos x 8 proposals x
9 segments RGB 224x224 (32 proposals, 288 frames), K=20 classes, STPP (1,(1,2),1), dropout 0,
frozen BN.  A step = BNInception fwd -> global-pool+STPP -> heads + multi-task loss (+ all
gradients) -> backbone bwd -> NCCL gradient allreduce (N>1) -> SGD step -> weight re-pack.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "action-detection_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

VIDEOS_PER_GPU, PROPS, SEG, K_CLASSES, STPP_CFG = 4, 8, 9, 20, (1, (1, 2), 1)
FLOP_PER_FRAME_FWD = 2 * 2031576064
FLOP_PER_FRAME_FWDBWD = 2 * (2031576064 + 1913562112 + 2031576064)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every ~5 ms; nvidia-smi as a fallback)."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]          # nvmlClocksEventReason*: HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap

    def __init__(self, index):
        self.rows, self.stop_flag, self.index = [], False, index      # rows: (sm_mhz, max_mhz, [reason flags])
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        self.rows.append((float(sm), float(mx), [bool(r & b) for b in self.BITS]))

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        c = [t.strip() for t in out.split(",")]
        if len(c) >= 6:
            self.rows.append((float(c[0]), float(c[1]), [t.lower().startswith("active") for t in c[2:6]]))

    def _run(self):
        while not self.stop_flag:
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.005 if self.nvml else 0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows]
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2][i] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml else "nvidia-smi"}


def usable_cores():
    """threads the host really grants this process: affinity mask capped by the cgroup CPU quota"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return n


def cpu_reference(steps, warmup, videos=2):
    """The reference's CPU PyTorch path for this workload, restated by the oracle (the reference
    scripts do not parse on py3.12), all host threads, bounded sample."""
    import torch
    from oracle import ssn_oracle as O, synth
    torch.set_num_threads(usable_cores())
    bb = synth.synth_backbone(3, seed=0, calib_frames=2)
    hd = synth.synth_heads(K_CLASSES, 5, seed=0)
    for d in (bb, hd):
        for k in d:
            if "_bn." not in k:
                d[k].requires_grad_(True)
    batch = synth.synth_batch(videos, K_CLASSES, 3, seed=0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        outs = O.ssn_train_forward(bb, hd, *batch, stpp_cfg=STPP_CFG)
        loss, _ = O.total_loss(outs)
        loss.backward()
        for d in (bb, hd):
            for v in d.values():
                v.grad = None
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    props = videos * PROPS
    best = min(times)
    mean = sum(times) / len(times)
    return {"value": props / mean, "best": props / best, "ms_per_step": mean * 1e3, "cores": torch.get_num_threads(),
            "sample": "%d videos = %d proposals x 9 seg fwd+bwd per step, %d steps" % (videos, props, len(times))}


WORKLOAD = ("THUMOS14-shape synthetic: batch 32 proposals x 9 segments RGB 224x224 per GPU, BNInception SSN "
            "fwd+bwd (+allreduce+SGD+repack), K=20, STPP (1,(1,2),1), dropout 0, frozen BN")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(max(1, min(args.steps, 2)), max(0, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": "proposals/sec (9-seg BNInception SSN fwd+bwd)", "value": r["value"],
            "unit": "proposals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch_proposals": 32 * args.gpus, "frames_per_gpu": 288,
                       "parallelism": "dp%d" % args.gpus, "precision": "f32 CPU",
                       "note": "reference CPU PyTorch path (oracle restatement, kind=port) timed on a bounded sample "
                               "(16 proposals per step) of the same workload; rank 0 only"},
            "cpu_baseline": {"value": r["value"], "unit": "proposals/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grad-scale", type=float, default=4096.0)
    ap.add_argument("--no-graph", action="store_true", help="do not capture the training step in a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import ssn_models
    from ssn_b200 import _lib
    from oracle import synth          # synthetic weights/inputs generator (test infrastructure, not measured)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.set_num_threads(max(1, usable_cores() // max(1, world)))     # host-side setup (synthetic data) shares the cores
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: whatever NCCL_DEBUG level the environment asks for goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/ssnb_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)

    prec = _lib.FAST_FP16 if args.precision == "fast" else _lib.EXACT_FP32
    torch.manual_seed(0)
    model = ssn_models.SSN(K_CLASSES, 2, 5, 2, "RGB", base_model="BNInception", dropout=0, stpp_cfg=STPP_CFG)
    bb = synth.synth_backbone(3, seed=0, calib_frames=2)
    sd = model.state_dict()
    for k, v in bb.items():
        sd["base_model." + k].copy_(v)
    model = model.to(dev).train()
    model.set_precision(prec, args.grad_scale)
    params = [p for p in model.parameters() if p.requires_grad]
    flat_grad = torch.zeros(sum(p.numel() for p in params), device=dev)
    off = 0
    for p in params:
        p.grad = flat_grad[off:off + p.numel()].view_as(p)
        off += p.numel()
    policies = model.get_optim_policies()
    groups = [{"params": g["params"], "lr": 1e-5 * g["lr_mult"], "weight_decay": 5e-4 * g["decay_mult"]} for g in policies if g["params"]]
    opt = torch.optim.SGD(groups, lr=1e-5, momentum=0.9)

    # per-rank shard of the global batch (weak scaling: fixed work per GPU); inputs resident in HBM
    nb = 2   # distinct device-resident batches, alternated
    batches = [tuple(t.to(dev) for t in synth.synth_batch(VIDEOS_PER_GPU, K_CLASSES, 3, seed=100 * rank + i)) for i in range(nb)]
    host_batches = [tuple(t.pin_memory() for t in synth.synth_batch(VIDEOS_PER_GPU, K_CLASSES, 3, seed=100 * rank + i)) for i in range(nb)]
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def eager_step(batch):
        flat_grad.zero_()
        losses = model.fused_step(*batch, global_videos=VIDEOS_PER_GPU * world, loss_scale=1.0 / world)
        if world > 1:
            dist.all_reduce(flat_grad)
        opt.step()
        return losses

    # The whole step (~420 kernel launches + all-reduce + optimizer + weight re-pack) is captured once in a
    # CUDA graph and replayed; the step's inputs are copied into the graph's static input tensors.
    step, used_graph = eager_step, False
    if not args.no_graph:
        try:
            static_batch = tuple(torch.empty_like(t) for t in batches[0])
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    for d_, s_ in zip(static_batch, batches[i % nb]):
                        d_.copy_(s_)
                    eager_step(static_batch)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_losses = eager_step(static_batch)

            def graph_step(batch):
                for d_, s_ in zip(static_batch, batch):
                    d_.copy_(s_)
                graph.replay()
                return static_losses
            step, used_graph = graph_step, True
        except Exception as ex:          # capture not possible on this stack: stay eager, say so
            if rank == 0:
                print("CUDA graph capture failed, running eagerly: %r" % (ex,), file=sys.stderr)
            torch.cuda.synchronize()
            step, used_graph = eager_step, False

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(batches[i % nb])
    barrier()
    launches0 = _lib.lib.ssnb_global_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    with ClockSampler(local) as clocks:
        for i in range(args.steps):
            l2_flush.zero_()                      # flush L2 between timed iterations (outside the event pair)
            ev[2 * i].record()
            losses = step(batches[i % nb])
            ev[2 * i + 1].record()
        barrier()
    launches = _lib.lib.ssnb_global_launch_count() - launches0
    if used_graph:      # replays do not pass through the library's counter: count one eager step and scale
        l0 = _lib.lib.ssnb_global_launch_count(); eager_step(batches[0]); torch.cuda.synchronize()
        launches = (_lib.lib.ssnb_global_launch_count() - l0) * args.steps
    ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps))
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    props_step = VIDEOS_PER_GPU * PROPS * world
    value = props_step * args.steps / (ms_total / 1e3)

    # ---- e2e: the reference-facing module calls (ssn_train.py:207-236) with HOST inputs ----------
    import ops.ssn_ops as R
    act_crit, comp_crit, reg_crit = torch.nn.CrossEntropyLoss(), R.CompletenessLoss(), R.ClassWiseRegressionLoss()

    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()

    def prefetch(hb):
        """host (pinned) -> device copy of one step's inputs on the copy stream; double-buffered so the
        copy of step i+1 overlaps the compute of step i (both inside the timed region)."""
        with torch.cuda.stream(copy_stream):
            db_ = tuple(t_.to(dev, non_blocking=True) for t_ in hb)
            evc = torch.cuda.Event()
            evc.record(copy_stream)
        return db_, evc

    def e2e_compute(db_, evc):
        main_stream.wait_event(evc)
        for t_ in db_:
            t_.record_stream(main_stream)
        x, sc, tg, rt, pt = db_
        flat_grad.zero_()
        a, at, c, ct, r, rl, rtt = model(x, sc, tg, rt, pt)
        loss = act_crit(a, at) + 0.1 * comp_crit(c, ct, 1, 7) + 0.1 * reg_crit(r, rl, rtt)
        (loss / world).backward()
        if world > 1:
            dist.all_reduce(flat_grad)
        opt.step()
        return loss

    def e2e_run(n):
        nxt = prefetch(host_batches[0])
        last = None
        for i in range(n):
            cur = nxt
            if i + 1 < n:
                nxt = prefetch(host_batches[(i + 1) % nb])
            loss = e2e_compute(*cur)
            if last is not None:
                last.item()                 # device -> host read of the previous step's result (one step of lag)
            last = loss
        return last.item()

    e2e_steps = max(3, args.steps // 2)
    e2e_run(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(e2e_steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = props_step * e2e_steps / (t.item() / 1e3)
    h2d = sum(t_.numel() * t_.element_size() for t_ in host_batches[0])

    # ---- roofline of the dominant kernel: per-op CUDA-event timing of the conv launches -----------
    peaks, peak_src = measured_peaks()
    roof = None
    if rank == 0:
        eng = model.base_model.engine_for(VIDEOS_PER_GPU * PROPS * SEG, True, dev)
        table = {n: (ci, co, k, s, p) for (n, ci, co, k, s, p) in __import__("ssn_b200.engine", fromlist=["conv_table"]).conv_table(3)}
        tot_ms, tot_flop, n_l = 0.0, 0.0, 0
        big = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i, (kind, iname, oname) in enumerate(eng.ops()):
            if kind != "conv":
                continue
            ci, co, k, s, p = table[oname[:-3]]
            if s != 1 or ci % 8:
                continue                           # conv1 / stride-2 layers use other plans of the same kernel
            _c, hh, ww = eng.value_shape(oname)
            best = 1e9
            for _ in range(3):
                l2_flush.zero_()
                a.record(); eng.run_op(i, False); b.record(); b.synchronize()
                best = min(best, a.elapsed_time(b))
            flop = 2.0 * eng.frames * hh * ww * co * ci * k * k
            tot_ms += best
            tot_flop += flop
            n_l += 1
            if big is None or flop > big[1]:
                big = (oname[:-3], flop, best)
        peak = peaks.get("bf16_tflops", 1590.0)
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("umma_conv_v2_kernel:%s_fwd" % big[0], {}).get("dram_bytes")
        except Exception:
            traffic = None
        achieved = big[1] / (big[2] / 1e3) / 1e12
        fam = tot_flop / (tot_ms / 1e3) / 1e12
        roof = {"bound": "tensor",
                "kernel": ("umma_conv_v2_kernel" if prec == _lib.FAST_FP16 else "conv_kernel<float> (SIMT)") + ", largest launch: %s forward" % big[0],
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_unit": "bytes per launch (ncu --set full dram read+write, profiles/r01_ncu_full_v2.txt)",
                "peak_source": peak_src + " bf16 burst (kernel timed alone, L2 flushed before each launch)",
                "flop_per_launch": big[1], "ms_per_launch": big[2],
                "family_average": {"launches_timed": n_l, "achieved": fam, "frac": fam / peak,
                                   "note": "all %d stride-1 forward launches of the kernel, sum(flop)/sum(time)" % n_l}}

    # ---- STPP bandwidth (the second half of BASELINE.json's metric): the fused global-pool + STPP kernel at the bench
    #      shape (28.9 MB of fp16 activations: launch-latency bound) and the standalone STPP kernel at a
    #      size where HBM bandwidth is the bound.  Reported beside the headline; a failure here never costs the bench line.
    stpp_info = None
    if rank == 0:
        try:
            import ctypes as C
            hbm = float(peaks.get("hbm_gbs", 6650.0))
            eng = model.base_model.engine_for(VIDEOS_PER_GPU * PROPS * SEG, True, dev)
            n_prop = VIDEOS_PER_GPU * PROPS
            lo, hi, nm, col = model.stpp.part_table([2, 7, 9])
            feat = torch.empty(n_prop * SEG, 1024, device=dev)
            course = torch.empty(n_prop, 1024, device=dev)
            pooled = torch.empty(n_prop, len(lo) * 1024, device=dev)
            sc = torch.rand(n_prop, 2, device=dev)
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(5):
                l2_flush.zero_()
                a.record()
                rc = _lib.lib.ssnb_gpool_stpp_fwd(eng.h, None, C.c_void_p(sc.data_ptr()), SEG, len(lo), _lib.int_array(lo), _lib.int_array(hi),
                                                  _lib.int_array(nm), _lib.int_array(col), 2, 7, C.c_void_p(feat.data_ptr()),
                                                  C.c_void_p(course.data_ptr()), C.c_void_p(pooled.data_ptr()), stream)
                b.record(); b.synchronize()
                if rc != 0:
                    raise RuntimeError("ssnb_gpool_stpp_fwd rc=%d" % rc)
                best = min(best, a.elapsed_time(b))
            fused_bytes = n_prop * SEG * 49 * 1024 * (2 if prec == _lib.FAST_FP16 else 4) + feat.numel() * 4 + course.numel() * 4 + pooled.numel() * 4
            big_n = 16384                                   # 16384 proposals x 9 segments x 1024 fp32 = 0.6 GB in, 0.4 GB out
            ft = torch.randn(big_n * SEG, 1024, device=dev)
            scb = torch.rand(big_n, 2, device=dev)
            big_best = 1e9
            for _ in range(4):
                l2_flush.zero_()
                a.record()
                ca, cc = model.stpp(ft, scb, [2, 7, 9])
                b.record(); b.synchronize()
                big_best = min(big_best, a.elapsed_time(b))
            big_bytes = ft.numel() * 4 + scb.numel() * 4 + ca.numel() * 4 + cc.numel() * 4
            stpp_info = {"fused_gpool_stpp": {"proposals": n_prop, "bytes": int(fused_bytes), "us": best * 1e3,
                                              "GB/s": fused_bytes / (best / 1e3) / 1e9, "frac_of_hbm_peak": fused_bytes / (best / 1e3) / 1e9 / hbm,
                                              "note": "bench shape; launch-latency bound at this size"},
                         "stpp_fwd_large": {"proposals": big_n, "bytes": int(big_bytes), "us": big_best * 1e3,
                                            "GB/s": big_bytes / (big_best / 1e3) / 1e9, "frac_of_hbm_peak": big_bytes / (big_best / 1e3) / 1e9 / hbm,
                                            "note": "StructuredTemporalPyramidPooling.forward, fp32, algorithmic bytes (61,448 B/proposal)"},
                         "hbm_peak_GB/s": hbm}
            del ft, scb, ca, cc
        except Exception as ex:                             # never lose the headline over the side measurement
            stpp_info = {"error": repr(ex)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(1, 1)
        cpu = {"value": r["value"], "unit": "proposals/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        line = {"metric": "proposals/sec (9-seg BNInception SSN fwd+bwd)", "value": value, "unit": "proposals/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate" if prec == _lib.FAST_FP16 else "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "global_batch_proposals": props_step, "frames_per_gpu": VIDEOS_PER_GPU * PROPS * SEG,
                           "parallelism": "dp%d" % world, "precision": args.precision, "l2": "flushed between timed steps (256 MiB write)",
                           "grad_scale": args.grad_scale, "cuda_graph": used_graph},
                "clocks": clocks.summary(), "gpu_launches": int(launches),
                "tflops_step": FLOP_PER_FRAME_FWDBWD * VIDEOS_PER_GPU * PROPS * SEG / (ms_total / args.steps / 1e3) / 1e12,
                "losses": [float(v) for v in losses.tolist()],
                "e2e": {"value": e2e_value, "unit": "proposals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                        "steps": e2e_steps, "path": "SSN.forward + CrossEntropy/CompletenessLoss/ClassWiseRegressionLoss + backward + SGD from pinned host tensors, H2D double-buffered on a copy stream, loss.item() every step"},
                "roofline": roof, "stpp": stpp_info, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        # all ranks leave together; skip NCCL/graph teardown (it can block when a captured graph holds the
        # communicator) — the process is done
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
