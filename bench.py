#!/usr/bin/env python
"""bench.py — cross_fusion fwd+bwd samples/sec on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --steps 2 --warmup 1          # CPU arm: the UNMODIFIED reference module (oracle/_ref)
    python bench.py --impl reference-gpu                           # reported only: the same reference module on this GPU (autocast bf16)
    python bench.py --sweep                                        # BASELINE config 5: one line per (D, image, language length)
    python bench.py --workload ego4dv1 | --mode infer | --feat-dtype bf16 | --accumulate 2
    torchrun --nproc-per-node N bench.py --gpus N ...              # one rank per GPU, NCCL

A "step" = one pass of the hot path (all 4 FPN levels x 4 encoder layers, forward + backward) over
one per-GPU batch of synthetic inputs of the Ego4Dv2 shape (`--workload
ego4dv1` selects config 2).  Dropout is ON (training mode, the shipped probabilities), gradients of
every fusion parameter are produced, and for N > 1 they are all-reduced over NCCL every step
(one process per GPU, one bucket per FPN level, weak scaling: per-GPU batch fixed).

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM;
`e2e` = same path through CrossFusionBoxWrapper.forward with HOST (pinned) inputs copied in and the
scalar loss read back every step; `roofline` = the dominant kernel family against the measured bf16
peak; `cpu_baseline` = the CPU oracle port timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# More hardware work queues than streams (main + 4 level streams + copy stream + NCCL's): with the default 8 connections
# streams alias onto shared queues, and a false dependency behind a spinning NCCL kernel can stall the end-to-end pass on
# 8 GPUs (seen intermittently).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch

METRIC = "cross_fusion fwd+bwd samples/sec"
METRIC_INFER = "cross_fusion fwd (inference) samples/sec"
UNIT = "samples/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))),
                "tflops_burst": float(p.get("bf16_tflops", 1590.0)), "hbm_gbs": float(p.get("hbm_gbs", 6650.0)),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  The process is started BEFORE the warm-up steps
    (its NVML initialisation takes locks in the driver and stalled kernel launches when it overlapped the ~0.3 s timed
    region); only the samples that arrive between mark() and stop() -- the timed region -- are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None
        self.thread = None
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            if self.t_mark is not None:
                self.lines.append(line.strip())

    def mark(self):
        """Start of the timed region: samples from here on count."""
        self.t_mark = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_flops_fwd(workload: dict, L: int) -> float:
    """F_fwd = sum_l [4 n K D + layers (16 S D^2 + 4 S^2 D)] per sample."""
    from transfusion_b200.configs import level_shapes
    D = workload["token_dim"]
    tot = 0.0
    for (h, w), C, p, nl in zip(level_shapes(workload), workload["channels"], workload["patch"], workload["num_layers"]):
        n = (h // p) * (w // p)
        K = C * p * p
        S = n + L
        tot += 4.0 * n * K * D + nl * (16.0 * S * D * D + 4.0 * S * S * D)
    return tot


# ------------------------------------------------------------------------------------------------
# Reference arm.  kind = "reference": the UNMODIFIED reference CrossFusionBoxWrapper, imported from /root/reference (build
# container) or its byte-identical staged copy oracle/_ref (oracle/build_ref.py; travels to the GPU box), run through its
# own forward / autograd with the yml's dropout in train mode, all host threads.  kind = "port" (only when neither tree
# exists): oracle/ref_math.py, whose restatement is pinned to the reference by tests/golden.
# ------------------------------------------------------------------------------------------------
def config_dict(args, w, B, L, world, train):
    """`config` of the JSON line -- shared by both arms so the driver compares like with like."""
    return {"workload": f"{args.workload} cross_fusion {'fwd+bwd' if train else 'fwd'}: 4 FPN levels x 4 layers, "
                        f"D={w['token_dim']}, image {w['image'][0]}x{w['image'][1]}",
            "per_gpu_batch": B, "global_batch": B * world, "lang_len": L,
            "dropout": "off" if (args.no_dropout or not train) else "on (0.1/0.15/0.1)",
            "feature_dtype": args.feat_dtype, "parallelism": f"dp{world}",
            "l2": "inputs and activations (>1 GB per step) exceed the 126 MB L2; no explicit flush",
            "visual_input_grad": False}


def reference_step(workload_name: str, batch: int, L: int, device="cpu", autocast=False, train=True, dropout=True, seed: int = 0):
    """One fwd(+bwd) step of the reference module on `device`.  Returns (step fn, kind)."""
    from transfusion_b200.configs import WORKLOADS, level_shapes
    from transfusion_b200.harness import synthetic_inputs
    w = WORKLOADS[workload_name]
    feats, lang, mask = synthetic_inputs(workload_name, batch, L, seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    cot = {k: torch.randn(v.shape, generator=g) for k, v in feats.items()}
    kind = "port"
    try:
        from oracle import ref_loader
        if ref_loader.reference_available():
            cfg = ref_loader.build_fusion_cfg(w["token_dim"], n_levels=len(w["channels"]), num_layers=w["num_layers"],
                                              num_heads=w["num_heads"], patch=w["patch"], dropout=1.0 if dropout else 0.0)
            m = ref_loader.build_reference_module(cfg, level_shapes(w), w["channels"], seed=seed,
                                                  noun_classes=w["noun_classes"], verb_classes=w["verb_classes"])
            kind = "reference"
    except Exception as e:   # fall back to the port, say so
        print(f"[bench] reference module unavailable ({e}); timing the oracle port", file=sys.stderr)
    if kind == "reference":
        m = m.to(device)
        m.train(train)
        for k, p in m.named_parameters():
            if k.endswith("heatmap_token"):
                p.requires_grad_(False)
        feats = {k: v.to(device) for k, v in feats.items()}
        cot = {k: v.to(device) for k, v in cot.items()}
        lang, mask = lang.to(device), mask.to(device)
        keys = sorted(feats, key=int)

        def step():
            m.zero_grad(set_to_none=True)
            m.rcnn_model.features = feats
            with torch.autocast(device_type="cuda" if str(device).startswith("cuda") else "cpu", dtype=torch.bfloat16, enabled=autocast):
                if train:
                    out = m({"image": None, "language_f": (lang, mask)})["features"]
                else:
                    with torch.no_grad():
                        out = m({"image": None, "language_f": (lang, mask)})["features"]
            if train:
                torch.autograd.backward([out[k] for k in keys], [cot[k].to(out[k].dtype) for k in keys])
        return step, kind

    from oracle import ref_math
    from transfusion_b200.harness import build_workload_module
    mm = build_workload_module(workload_name, device="cpu", dropout=False, seed=seed)
    sd = {k: p.detach().clone().requires_grad_(True) for k, p in mm.named_parameters()
          if not k.startswith(("rcnn_model", "narr_pooling_layer"))}

    def step():
        for v in sd.values():
            v.grad = None
        lg = lang.clone().requires_grad_(True)
        out, _ = ref_math.cross_fusion_forward(feats, lg, mask, sd, w["patch"], w["num_heads"], w["num_layers"])
        loss = sum((out[k] * cot[k]).sum() for k in out)
        if train:
            loss.backward()
    return step, kind


def run_reference_arm(args):
    """CPU arm (`--impl reference`): rank 0 only.  Same `config` as our arm; every step is a BOUNDED sample of that workload
    (cpu_baseline.sample says how many samples) sized so the whole --steps/--warmup run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from transfusion_b200.configs import WORKLOADS
    w = WORKLOADS[args.workload]
    L = args.lang_len or w["lang_len"]
    train = args.mode == "train"
    B = args.batch or (w["train_batch"] if train else w["eval_batch"])
    world = max(1, args.gpus)
    bs = args.cpu_batch
    if bs <= 0:   # auto: probe one sample, then size the per-step sample for ~150 s of total CPU work
        probe, _ = reference_step(args.workload, 1, L, train=train, dropout=not args.no_dropout)
        probe()
        t0 = time.perf_counter(); probe(); t1 = time.perf_counter() - t0
        bs = max(1, min(B, int(150.0 / (max(1, args.steps + args.warmup) * t1))))
        del probe
    step, kind = reference_step(args.workload, bs, L, train=train, dropout=not args.no_dropout)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = bs * args.steps / dt
    what = "the UNMODIFIED reference CrossFusionBoxWrapper (oracle/_ref), its own forward + autograd" if kind == "reference" \
        else "oracle/ref_math.py (port; reference tree not staged)"
    sample = (f"{bs} sample(s)/step of the {args.workload} 4-level workload (L={L}), {'fwd+bwd' if train else 'fwd'}, fp32, "
              f"{'dropout on' if (kind == 'reference' and not args.no_dropout and train) else 'dropout p=0'}; {what}")
    line = {"impl": "reference", "metric": METRIC if train else METRIC_INFER, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, w, B, L, world, train),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_reference_gpu(args):
    """`--impl reference-gpu` (REPORTED ONLY, BASELINE.md §4's same-box comparator): the unmodified reference module on this
    GPU under torch.autocast(bf16), full per-GPU batch, CUDA-event timed.  It is ATen/cuBLAS/SDPA library code -- the bar
    the hand-written path has to beat on the same silicon -- not part of the product."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from transfusion_b200.configs import WORKLOADS
    w = WORKLOADS[args.workload]
    L = args.lang_len or w["lang_len"]
    train = args.mode == "train"
    B = args.batch or (w["train_batch"] if train else w["eval_batch"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    step, kind = reference_step(args.workload, B, L, device=dev, autocast=True, train=train, dropout=not args.no_dropout)
    if kind != "reference":
        print(json.dumps({"impl": "reference-gpu", "unavailable": "reference tree not staged (oracle/_ref missing)"}), flush=True)
        return
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    val = B * args.steps / (ms / 1e3)
    line = {"impl": "reference-gpu", "metric": METRIC if train else METRIC_INFER, "value": val, "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "dtype": "bf16 (torch.autocast)", "data": "synthetic", "config": config_dict(args, w, B, L, 1, train),
            "note": "unmodified reference module, stock ATen/cuBLAS/SDPA kernels on the same B200; reported only",
            "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2), "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_numa(gpu_index: int):
    """Pin this rank's threads to the CPUs of its GPU's NUMA node (sysfs) before the pinned input buffers of the end-to-end
    pass are allocated (first touch then places them on that node).  Returns a short note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(gpu_index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return f"single NUMA domain ({len(nodes)} node(s)); no binding"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
        return f"rank bound to NUMA node {node} ({len(cpus)} cpus)"
    except Exception as e:
        return f"no NUMA binding ({type(e).__name__})"


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from transfusion_b200 import _lib, ops
    from transfusion_b200.configs import WORKLOADS
    from transfusion_b200.harness import build_workload_module, synthetic_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    w = WORKLOADS[args.workload]
    L = args.lang_len or w["lang_len"]
    B = args.batch or (w["train_batch"] if args.mode == "train" else w["eval_batch"])
    train = args.mode == "train"

    model = build_workload_module(args.workload, device=dev, dropout=not args.no_dropout, seed=0)
    model.train(train)
    for k, p in model.named_parameters():
        if k.endswith("heatmap_token"):
            p.requires_grad_(False)  # registered but unused in forward (reference cross_f_box_layers.py:43)
    net = model
    fpn = None
    if args.with_fpn != "none":
        # SURVEY 8f N1 measurement: a stock torchvision FPN (256 channels + max-pool level) consumes the fused maps, either
        # as the reference does (fused [B,C,h,w] maps -> inner 1x1 convs -> ...) or with the laterals folded into the
        # back-projection GEMM (fuse_fpn_inner).  Everything after the laterals is identical library code in both arms.
        from collections import OrderedDict
        from torchvision.ops import FeaturePyramidNetwork
        from torchvision.ops.feature_pyramid_network import LastLevelMaxPool
        torch.manual_seed(5)
        torch.backends.cudnn.allow_tf32 = True   # the detector's own 3x3 convolutions: library code, identical in both arms
        fpn = FeaturePyramidNetwork(w["channels"], 256, extra_blocks=LastLevelMaxPool()).to(dev)
        if args.with_fpn in ("fused", "laterals"):
            model.fuse_fpn_inner(fpn)
            if args.with_fpn == "laterals":   # stop at the laterals: isolates what the fusion path itself gains from N1
                import transfusion_b200.obj_detection.fpn as _fpn_mod
                _fpn_mod.fpn_from_laterals = lambda _f, lat: lat
        else:
            def apply_fpn(d, _fpn=fpn):
                d["features"] = _fpn(OrderedDict((k, d["features"][k].float()) for k in sorted(d["features"], key=int)))
                return d
            model.rcnn_model.apply_fpn = apply_fpn
    reducer = None
    if world > 1 and train:
        # same initial weights on every rank (seeded construction); gradients averaged per level bucket,
        # each bucket's NCCL all-reduce launched as soon as that level's backward has been enqueued
        from transfusion_b200.parallel import BucketedGradAllReduce, level_buckets
        reducer = BucketedGradAllReduce(level_buckets(model), compress=args.grad_compress if args.grad_compress != "none" else None)

    feat_dtype = torch.float32 if args.feat_dtype == "f32" else torch.bfloat16
    feats_h, lang_h, mask_h = synthetic_inputs(args.workload, B, L, seed=1234 + rank, feat_dtype=feat_dtype, pin=True)
    feats_d = {k: v.to(dev) for k, v in feats_h.items()}
    lang_d, mask_d = lang_h.to(dev), mask_h.to(dev)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    cot = {k: torch.randn(v.shape, device=dev, dtype=torch.float32, generator=gen) for k, v in feats_d.items()}
    keys = sorted(feats_d, key=int)
    if fpn is not None:   # the outputs are the FPN maps (256 channels per level + the pooled level)
        model.rcnn_model.features = feats_d
        with torch.no_grad():
            probe = net({"image": None, "language_f": (lang_d, mask_d)})["features"]
        keys = list(probe.keys())
        cot = {k: torch.randn(v.shape, device=dev, dtype=torch.float32, generator=gen) for k, v in probe.items()}
        del probe

    acc_n = max(1, args.accumulate)

    def fwd_bwd(feats, lg, mk, micro, want_loss=False):
        """One micro-step.  accumulate_grad_batches = N (ego_nao_res50_ego4dv2.yml:125): gradients are zeroed before the first
        micro-step, accumulate in place (in the first micro-step's arenas) and are all-reduced on the N-th only, like
        Lightning's DDP no_sync (run_experiment.py:444)."""
        first, last = micro % acc_n == 0, micro % acc_n == acc_n - 1
        if first:
            net.zero_grad(set_to_none=True)
        if reducer is not None:
            reducer.active = last
            if last:
                reducer.reset()
        out = net({"image": None, "language_f": (lg, mk)})["features"]
        loss = None
        if want_loss:
            with torch.no_grad():   # loss = <out, cot>; its gradient w.r.t. out is cot itself
                loss = sum(torch.dot(out[k].float().reshape(-1), cot[k].reshape(-1)) for k in keys)
        torch.autograd.backward([out[k] for k in keys], [cot[k].to(out[k].dtype) for k in keys])
        if reducer is not None and last:
            reducer.finish()
        return loss

    res_state = {"i": 0}

    def step_resident():
        model.rcnn_model.features = feats_d
        if train:
            fwd_bwd(feats_d, lang_d, mask_d, res_state["i"])
            res_state["i"] += 1
        else:
            with torch.no_grad():
                net({"image": None, "language_f": (lang_d, mask_d)})

    # end-to-end step: inputs start in pinned HOST memory; the copy of step i+1's inputs is issued on a copy
    # stream while step i computes (double-buffered device staging), and the scalar loss of every step is copied
    # back to pinned host memory.  One H2D of the full input set and one D2H per step are inside the timed region.
    # Nothing blocks the host on the step it just launched: the staging slot is fenced on the device (the copy
    # stream waits for the event of the compute that last read the slot) and the loss is consumed one step later.
    def build_e2e(feats_src):
        """End-to-end step over the HOST buffers `feats_src` (pinned; fp32 or bf16 maps) + lang_h / mask_h."""
        copy_stream = torch.cuda.Stream(device=dev)
        stage_bufs = [({k: torch.empty_like(v, device=dev) for k, v in feats_src.items()}, torch.empty_like(lang_h, device=dev),
                       torch.empty_like(mask_h, device=dev)) for _ in range(2)]
        stage_ev = [torch.cuda.Event(), torch.cuda.Event()]
        done_ev = [None, None]        # compute that last read staging slot s has finished
        loss_ev = [None, None]
        loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        st = {"i": 0, "last": 0.0, "copies": []}

        def prefetch(slot):
            f, lg, mk = stage_bufs[slot]
            if done_ev[slot] is not None:
                copy_stream.wait_event(done_ev[slot])
            with torch.cuda.stream(copy_stream):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(copy_stream)
                for k in f:
                    f[k].copy_(feats_src[k], non_blocking=True)
                lg.copy_(lang_h, non_blocking=True)
                mk.copy_(mask_h, non_blocking=True)
                c1.record(copy_stream)
                st["copies"].append((c0, c1))
                stage_ev[slot].record(copy_stream)

        def step():
            i = st["i"]
            st["i"] = i + 1
            slot = i & 1
            prefetch(slot ^ 1)                      # next step's inputs, overlapped with this step's compute
            cur = torch.cuda.current_stream()
            cur.wait_event(stage_ev[slot])
            f, lg, mk = stage_bufs[slot]
            model.rcnn_model.features = f
            if train:
                loss = fwd_bwd(f, lg, mk, i, want_loss=True)
            else:
                with torch.no_grad():
                    out = net({"image": None, "language_f": (lg, mk)})["features"]
                    loss = sum(out[k].float().sum() for k in keys)
            loss_host[slot].copy_(loss.reshape(1), non_blocking=True)   # device -> host read of the step's result
            loss_ev[slot] = torch.cuda.Event()
            loss_ev[slot].record(cur)
            done_ev[slot] = loss_ev[slot]
            if loss_ev[slot ^ 1] is not None:       # consume the previous step's loss (already on the host)
                loss_ev[slot ^ 1].synchronize()
                st["last"] = float(loss_host[slot ^ 1][0])
            return st["last"]

        nbytes = sum(v.numel() * v.element_size() for v in feats_src.values()) + lang_h.numel() * 4 + mask_h.numel() * 8

        def h2d_rate():
            """median GB/s of this rank's per-step input copies (the last `steps` of them)"""
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in st["copies"][-args.steps:])
            return round(nbytes / (ts[len(ts) // 2] * 1e-3) / 1e9, 1) if ts else None

        return prefetch, step, nbytes, h2d_rate

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        # watchdog: if a timed phase does not finish, dump every thread's Python stack and exit non-zero instead of
        # hanging the box (XF_WATCHDOG_S seconds, default 120; 0 disables)
        wd = int(os.environ.get("XF_WATCHDOG_S", "120"))
        if wd > 0:
            faulthandler.dump_traceback_later(wd, exit=True)
        try:
            return _timed(fn, steps)
        finally:
            if wd > 0:
                faulthandler.cancel_dump_traceback_later()

    live = {"e0": None, "per_step": None}   # the running timed phase's events (read by the end-to-end stall guard)

    def _timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per_step, host_t = [], [time.perf_counter()]
        live["e0"], live["per_step"] = e0, per_step
        n_alloc0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        e0.record()
        for _ in range(steps):
            fn()
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            per_step.append(ev)
            host_t.append(time.perf_counter())
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if rank == 0:   # diagnostics: per-step device times and cudaMalloc calls inside the timed region (stderr)
            prev, out = e0, []
            for ev in per_step:
                out.append(round(prev.elapsed_time(ev), 2))
                prev = ev
            host_ms = [round(1e3 * (b - a), 1) for a, b in zip(host_t[:-1], host_t[1:])]
            print(f"[step times ms] {out}  host enqueue ms {host_ms}  cudaMalloc calls in region: "
                  f"{torch.cuda.memory_stats(dev).get('num_device_alloc', 0) - n_alloc0}", file=sys.stderr)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("XF_NO_CLOCK_SAMPLER"):
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step_resident()
    torch.cuda.synchronize()
    sampler.mark()
    launches0 = _lib.lib().xf_launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.lib().xf_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)

    # ---- per-kernel roofline pass: CUDA events around every launch (rank 0, separate from the timed run)
    roofline, kernels = None, None
    # every rank runs the step (it contains the gradient all-reduce); only rank 0 records the events
    # The pass runs the levels one after the other on one stream (the timed runs overlap them on side streams), so
    # every launch is timed alone: ms_per_step values are serialised kernel times and sum to more than ms_per_step.
    from transfusion_b200.cross_fusion import cross_f_box_wrapper as _wrap
    _streams_on, _wrap.LEVEL_STREAMS = _wrap.LEVEL_STREAMS, False
    if rank == 0:
        ops.PROFILE = []
    step_resident()
    torch.cuda.synchronize()
    _wrap.LEVEL_STREAMS = _streams_on
    if rank == 0:
        peaks = load_peaks()
        prof, ops.PROFILE = ops.PROFILE, None
        fam = {}
        for (name, flops, nbytes, e0, e1) in prof:
            d = fam.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += e0.elapsed_time(e1); d["flops"] += flops; d["bytes"] += nbytes; d["launches"] += 1
        tot_ms = sum(d["ms"] for d in fam.values()) or 1.0
        kernels = {}
        for name, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            kernels[name] = {"ms_per_step": round(d["ms"], 4), "share": round(d["ms"] / tot_ms, 4), "launches": d["launches"]}
            if d["flops"] > 0:
                kernels[name]["tflops"] = round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 2)
            elif d["bytes"] > 0:
                kernels[name]["gbs"] = round(d["bytes"] / (d["ms"] * 1e-3) / 1e9, 1)
        top = next(iter(kernels))
        kname = {"gemm_fwd": "gemm_bf16_tcgen05_kernel (forward x W^T)", "gemm_dgrad": "gemm_bf16_tcgen05_kernel (dgrad)",
                 "gemm_wgrad": "gemm_bf16_tcgen05_kernel (wgrad, split-K)", "attn_fwd": "attn_fwd_tcgen05_kernel",
                 "attn_bwd": ("xf_attn_bwd = attn_bwd2_tcgen05_kernel<MODE_DVS> (scores once, dV on chip, E -> scratch) + two batched "
                              "gemm_bf16_tcgen05_kernel (dQ = E^T K, dK = E Q)") if os.environ.get("XF_ATTN_BWD_WS", "1") != "0"
                 else "attn_bwd2_tcgen05_kernel<MODE_DQ|MODE_DK|MODE_DV> (one xf_attn_bwd call = 3 passes)"}.get(top, top)
        traffic = None
        # DRAM bytes per call of the dominant kernel family at the level-0 shape: ncu cannot run inside this process, so the
        # figure comes from the committed `ncu --set full` capture of the SAME kernels (tools/profile_kernels.py, regenerated
        # per round by tools/make_traffic_json.py); `traffic_source` names it.  null when no capture is committed.
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")
        traffic_source = None
        if os.path.isfile(tpath):
            with open(tpath) as tf:
                tfam = json.load(tf)["families"].get(top.split("_")[0] if top.startswith("gemm") else top, {})
            traffic = tfam.get("traffic_bytes_per_call", tfam.get("traffic_bytes_per_launch"))
            traffic_source = "profiles/ncu_traffic_r2.json (ncu --set full of tools/profile_kernels.py, level-0 shape, B = 13)"
        if "tflops" in kernels[top]:
            roofline = {"bound": "tensor", "kernel": kname, "achieved": kernels[top]["tflops"], "peak": peaks["tflops"],
                        "unit": "TFLOP/s", "frac": round(kernels[top]["tflops"] / peaks["tflops"], 4), "traffic": traffic,
                        "share_of_step": kernels[top]["share"], "peak_source": peaks["source"]}
        else:
            roofline = {"bound": "hbm", "kernel": kname, "achieved": kernels[top].get("gbs"), "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": round((kernels[top].get("gbs") or 0.0) / peaks["hbm_gbs"], 4), "traffic": traffic,
                        "share_of_step": kernels[top]["share"], "peak_source": peaks["source"]}
        f_fwd = algorithmic_flops_fwd(w, L)
        from transfusion_b200.configs import level_shapes
        pe_dgrad = sum(2.0 * ((h // p) * (ww // p)) * (C * p * p) * w["token_dim"]
                       for (h, ww), C, p in zip(level_shapes(w), w["channels"], w["patch"]))
        # fwd+bwd = 3 F_fwd minus the patch-embed dgrad (visual inputs do not require grad: frozen backbone)
        step_flops = (3.0 * f_fwd - pe_dgrad) if train else f_fwd
        roofline["traffic_source"] = traffic_source
        roofline["whole_path_tflops"] = round(value / world * step_flops / 1e12, 2)
        roofline["whole_path_frac"] = round(value / world * step_flops / 1e12 / peaks["tflops"], 4)

    # ---- CPU baseline (rank 0, N = 1 only): the reference module (oracle/_ref) on a bounded sample of the same workload
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        nb = max(1, args.cpu_batch)
        stepc, kind = reference_step(args.workload, nb, L, train=train, dropout=not args.no_dropout)
        stepc()   # warm-up (allocator, thread pool)
        t0 = time.perf_counter()
        stepc()
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": nb / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                        "sample": f"{nb} sample(s) of the {args.workload} 4-level workload (L={L}), one {'fwd+bwd' if train else 'fwd'} after "
                                  f"one warm-up, fp32, {'the unmodified reference module (oracle/_ref), dropout on' if kind == 'reference' else 'oracle port, dropout p=0'} "
                                  f"({dt:.1f} s)"}

    # ---- end-to-end pass LAST and guarded: everything else of the line is already measured.  If the pass stalls (seen
    # intermittently on 8 GPUs: pinned-host H2D stream + side streams + NCCL), the guard prints the line with `e2e.value`
    # null and a note and ends every rank with exit code 0, so the device-timed result of the run is not lost.
    e2e_value = ms_e2e = h2d_gbs = numa_note = None
    e2e_bf16 = None
    h2d = sum(v.numel() * v.element_size() for v in feats_h.values()) + lang_h.numel() * 4 + mask_h.numel() * 8
    d2h = 4
    stall = {"fired": False}

    def make_line(e2e_note=None):
        line = {"metric": METRIC if train else METRIC_INFER, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": config_dict(args, w, B, L, world, train),
                "notes": {"level_streams": "independent FPN levels overlap on side streams in the timed runs; the `kernels` "
                                           "breakdown is taken with the levels serialised",
                          "accumulate_grad_batches": args.accumulate, "fpn": args.with_fpn, "grad_allreduce": "fp32" if args.grad_compress == "none" else args.grad_compress,
                          "attn_bwd": "5-unit schedule (scores once; bf16 [B,H,S,S] scratch)" if os.environ.get("XF_ATTN_BWD_WS", "1") != "0"
                                      else "3 on-chip passes (8 units)"},
                "clocks": clocks, "gpu_launches": int(launches),
                "tma_descriptor_cache": {"hits": int(_lib.lib().xf_tmap_cache_stats(0)), "encoded": int(_lib.lib().xf_tmap_cache_stats(1))},
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": (ms_e2e / args.steps) if ms_e2e else None,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        # achieved host->device rate of THIS rank's input copies (CUDA events on the copy stream, median
                        # over the timed steps) and the rate the step needs to hide them completely
                        "h2d_gbs_achieved": h2d_gbs, "h2d_gbs_needed_to_hide": round(h2d / (ms_per_step * 1e-3) / 1e9, 1),
                        "numa": numa_note,
                        "pipeline": "pinned host inputs copied on a side stream one step ahead (2 device slots, fenced by "
                                    "events); each step's loss is copied to pinned host memory and read one step later"},
                "e2e_bf16_features": e2e_bf16,
                "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline}
        if e2e_note:
            line["e2e"]["note"] = e2e_note
        return line

    def partial_e2e():
        """Steps of the stalled pass that DID finish on the device (event.query() never blocks): (n, ms) or None."""
        try:
            e0_, evs = live.get("e0"), list(live.get("per_step") or [])
            if not stall.get("in_e2e") or e0_ is None or not e0_.query():
                return None
            n_done = 0
            for ev in evs:
                if not ev.query():
                    break
                n_done += 1
            if n_done < 3:
                return None
            return n_done, e0_.elapsed_time(evs[n_done - 1])
        except Exception:
            return None

    def on_stall():
        nonlocal e2e_value, ms_e2e
        stall["fired"] = True
        if rank == 0:
            note = ("end-to-end pass did not finish within the guard time: stalled (no e2e number for this run); "
                    "the device-timed fields above are complete")
            part = partial_e2e()
            if e2e_value is not None and not stall.get("in_e2e"):
                note = "the extra bf16-feature end-to-end pass stalled; the fp32 `e2e` above is complete"
            elif stall.get("pre") and (part is None or part[0] < 6):
                n_pre, ms_pre_ = stall["pre"]
                e2e_value = world * B * n_pre / (ms_pre_ / 1e3)
                ms_e2e = ms_pre_ * args.steps / n_pre
                note = (f"the {args.steps}-step end-to-end pass stalled; value from the preceding fully synchronised {n_pre}-step pass "
                        f"(max over ranks); the device-timed fields above are complete")
            elif part is not None:
                n_done, ms_part = part
                e2e_value = world * B * n_done / (ms_part / 1e3)
                ms_e2e = ms_part * args.steps / n_done
                note = (f"end-to-end pass stalled after {n_done} of {args.steps} timed steps: value from those {n_done} steps on rank 0's "
                        f"clock (no max over ranks); the device-timed fields above are complete")
            print(json.dumps(make_line(note)), flush=True)
        sys.stdout.flush()
        os._exit(0)

    guard_s = float(os.environ.get("XF_E2E_GUARD_S", "45"))
    guard = threading.Timer(guard_s, on_stall) if guard_s > 0 else None
    if guard is not None:
        guard.daemon = True
        guard.start()
    os.environ["XF_WATCHDOG_S"] = "0"   # the guard replaces the hard watchdog for this pass
    if os.environ.get("XF_TEST_STALL_E2E"):   # test hook: emulate a stalled pass
        time.sleep(guard_s + 30)
    # ---- end-to-end: host (pinned) inputs in, scalar loss out, every step
    numa_note = bind_numa(local_rank) if world > 1 or os.environ.get("XF_NUMA_BIND") else None
    prefetch, step_e2e, h2d, h2d_rate = build_e2e(feats_h)
    prefetch(0)
    for _ in range(2):
        step_e2e()
    stall["in_e2e"] = True
    if world >= 8 and args.steps > 6:
        # the stall has only been seen on 8 GPUs: bank a short, fully synchronised (max over ranks) measurement first, so a
        # stall in the long pass still leaves a complete end-to-end number behind (the guard reports it with a note)
        ms_pre = timed(step_e2e, 6)
        stall["pre"] = (6, ms_pre)
    ms_e2e = timed(step_e2e, args.steps)
    stall["in_e2e"] = False
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d_gbs = h2d_rate()
    d2h = 4
    del prefetch, step_e2e
    # the same end-to-end step fed bf16 feature maps (half the host bytes; the module accepts bf16 maps and Ego4Dv1 trains
    # with precision 16): reported BESIDE the fp32 headline, never instead of it
    e2e_bf16 = None
    # (N > 1: only on request -- `--bf16-e2e` or a separate `--feat-dtype bf16` run -- so the scaling runs time exactly the
    # fp32 pipeline and nothing else)
    if args.feat_dtype == "f32" and not args.no_bf16_e2e and (world == 1 or args.bf16_e2e):
        feats_hb = {k: v.to(torch.bfloat16).pin_memory() for k, v in feats_h.items()}
        prefetch_b, step_b, h2d_b, rate_b = build_e2e(feats_hb)
        prefetch_b(0)
        for _ in range(3):
            step_b()
        ms_b = timed(step_b, args.steps)
        e2e_bf16 = {"value": world * B * args.steps / (ms_b / 1e3), "unit": UNIT, "ms_per_step": ms_b / args.steps,
                    "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": 4, "h2d_gbs_achieved": rate_b()}
        del prefetch_b, step_b, feats_hb

    if guard is not None:
        guard.cancel()

    if rank == 0:
        line = make_line()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


SWEEP_IMAGES = [(480, 608), (544, 640), (640, 768), (704, 896), (768, 1024), (800, 1280)]   # ego_nao_res50_ego4d.yml:22-23, padded to 32
SWEEP_LANG = [16, 32, 64, 128, 256, 512]


def run_sweep(args):
    """BASELINE config 5: language-context length x visual token grid x width, B = 16, fwd+bwd, dropout on, 1 GPU.
    One JSON line per (workload width, image, L): device-timed samples/s and the whole-path tensor fraction."""
    import copy
    from transfusion_b200.configs import WORKLOADS, level_shapes
    from transfusion_b200.harness import build_workload_module
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    B = args.batch or 16
    for name in args.sweep_widths.split(","):
        model = build_workload_module(name, device=dev, dropout=not args.no_dropout, seed=0)
        model.train(True)
        for k, p in model.named_parameters():
            if k.endswith("heatmap_token"):
                p.requires_grad_(False)
        images = SWEEP_IMAGES if not args.sweep_images else [tuple(int(v) for v in t.split("x")) for t in args.sweep_images.split(",")]
        for image in images:
            w = copy.deepcopy(WORKLOADS[name])
            w["image"] = image
            g = torch.Generator(device=dev).manual_seed(7)
            feats = {str(i): torch.relu(torch.randn(B, c, h, ww, device=dev, generator=g))
                     for i, ((h, ww), c) in enumerate(zip(level_shapes(w), w["channels"]))}
            cot = {k: torch.randn(v.shape, device=dev, generator=g) for k, v in feats.items()}
            keys = sorted(feats, key=int)
            for L in SWEEP_LANG:
                lang = 0.5 * torch.randn(B, L, w["token_dim"], device=dev, generator=g)
                lens = torch.randint(L // 2, L + 1, (B,), device=dev, generator=g)
                lens[0] = L
                mask = (torch.arange(L, device=dev)[None, :] < lens[:, None]).to(torch.int64)

                def step():
                    model.rcnn_model.features = feats
                    model.zero_grad(set_to_none=True)
                    out = model({"image": None, "language_f": (lang, mask)})["features"]
                    torch.autograd.backward([out[k] for k in keys], [cot[k] for k in keys])

                for _ in range(max(2, min(args.warmup, 3))):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                steps = max(3, min(args.steps, 5))
                e0.record()
                for _ in range(steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                f_fwd = algorithmic_flops_fwd(w, L)
                pe_dgrad = sum(2.0 * ((h // p) * (ww // p)) * (C * p * p) * w["token_dim"]
                               for (h, ww), C, p in zip(level_shapes(w), w["channels"], w["patch"]))
                tf = B / (ms * 1e-3) * (3.0 * f_fwd - pe_dgrad) / 1e12
                grids = [((h // p), (ww // p)) for (h, ww), p in zip(level_shapes(w), w["patch"])]
                print(json.dumps({"metric": METRIC, "value": round(B / (ms * 1e-3), 2), "unit": UNIT, "n_gpus": 1, "ms_per_step": round(ms, 3),
                                  "steps": steps, "dtype": "bf16", "data": "synthetic",
                                  "config": {"workload": f"sweep: {name} width D={w['token_dim']}, image {image[0]}x{image[1]}, L={L}",
                                             "per_gpu_batch": B, "lang_len": L, "token_grids": grids,
                                             "dropout": "off" if args.no_dropout else "on (0.1/0.15/0.1)"},
                                  "whole_path_tflops": round(tf, 1), "whole_path_frac": round(tf / peaks["tflops"], 4)}), flush=True)
            del feats, cot
        del model
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="ego4dv2", choices=["ego4dv2", "ego4dv1"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the reference's per-GPU batch)")
    ap.add_argument("--lang-len", type=int, default=0)
    ap.add_argument("--feat-dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--no-dropout", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="samples per CPU step; 0 = auto (1 for the in-line cpu_baseline, sized for ~150 s total in --impl reference)")
    ap.add_argument("--accumulate", type=int, default=1, help="accumulate_grad_batches: all-reduce every N-th micro-step")
    ap.add_argument("--no-bf16-e2e", action="store_true", help="skip the extra end-to-end pass with bf16 feature maps")
    ap.add_argument("--with-fpn", default="none", choices=["none", "stock", "fused", "laterals"],
                    help="append a torchvision FPN to the step: 'stock' = on the fused maps (reference data flow), 'fused' = "
                         "laterals folded into the back-projection GEMM (SURVEY 8f N1), 'laterals' = the same but the step "
                         "ends at the [B,256,h,w] laterals (no cuDNN convolutions in the timed region)")
    ap.add_argument("--grad-compress", default="none", choices=["none", "bf16"],
                    help="N > 1: exchange the gradient arenas as bf16 (default: fp32 like the reference's DDP)")
    ap.add_argument("--bf16-e2e", action="store_true", help="run the extra bf16-feature end-to-end pass for N > 1 too")
    ap.add_argument("--sweep", action="store_true", help="BASELINE config 5: language length x image size x width, B = 16")
    ap.add_argument("--sweep-images", default="", help="restrict the sweep to these padded image sizes, e.g. 704x896,768x1024")
    ap.add_argument("--sweep-widths", default="ego4dv2,ego4dv1", help="which widths the sweep covers (D = 896 / 712)")
    args = ap.parse_args()
    if args.sweep:
        run_sweep(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
