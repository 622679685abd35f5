#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MissM-Benchmark hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): image + depth + thermal ViT-L/14 224px towers + `sum` fusion
head, fwd+bwd, bf16 tensor-core GEMMs with fp32 accumulate / fp32 residual stream, B = 64 samples
per GPU, 30 % of the samples missing one modality (codes drawn as src/utils/generate_missing.py
does), synthetic inputs and name-seeded synthetic weights.  A step = loss.backward() through
finetune_model.forward (src/model/baseline.py:450-453) exactly as train_ddp.py:249-253 calls it; at
N > 1 the model is wrapped in torch DDP (train_ddp.py:189) so gradients are bucket-allreduced with
NCCL during backward.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "train samples/sec (fwd+bwd)"
MODALS = ['image', 'depth', 'thermal']
WORKLOAD = "configs[1]: image+depth+thermal ViT-L/14 towers + sum fusion head, fwd+bwd bf16, B=64/GPU, 30% missing"
FWD_GFLOP = {'image': 162.0, 'depth': 162.0, 'thermal': 162.0, 'audio': 393.4, 'video': 1711.7, 'language': 13.3}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="missm", choices=["missm", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU")
    ap.add_argument("--missing", type=float, default=0.3)
    ap.add_argument("--cpu-baseline-samples", type=int, default=0,
                    help="samples per CPU step (0 = auto: 2 for the cpu_baseline leg, 2..8 for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--layers", type=int, default=24, help=argparse.SUPPRESS)  # debugging only
    # variants beyond the headline workload; the defaults ARE the headline
    ap.add_argument("--lora-r", type=int, default=0, help="peft-style LoRA rank on the towers' attention projections "
                    "(reference config default 2; encoder frozen, adapters trained)")
    ap.add_argument("--optimizer", default="none", choices=["none", "fused", "torch"],
                    help="also time optimizer.step(): missm_b200.optim.FusedAdam or torch.optim.Adam")
    return ap.parse_args()


def full_configs(layers=24, lora_r=0):
    import restatement as R
    from missm_b200 import config as C
    v = {k: val for k, val in C.VIT_L14.items() if k != 'lora_r'}
    v['num_hidden_layers'] = layers
    if lora_r:
        v['lora_r'], v['lora_alpha'] = lora_r, 16
    t = dict(C.CLIP_TEXT)
    cfgs = {m: R.vision_config(**v) for m in MODALS}
    return cfgs, R.text_config(**t)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/restatement.py) of the reference path on the host cores
# ------------------------------------------------------------------------------------------------
def reference_step_fn(n_samples, layers=24):
    """One fwd+bwd of the workload on `n_samples` samples through the UNMODIFIED reference: its own LanguageBind
    bank (languagebind/__init__.py:54-85), finetune_model + `sum` head (src/model/baseline.py:43-61, 421-453) and
    CLIP towers, imported from /root/reference (build container) or its staged copy oracle/_ref (GPU box) through
    oracle/ref_shim.py, fp32 on the host cores, full batch through every tower (the reference never skips missing
    samples).  Returns None when neither tree is present."""
    import torch
    import ref_shim
    import restatement as R
    from missm_b200 import config as C
    if not ref_shim.available():
        return None
    v = dict(C.VIT_L14, lora_r=0, num_hidden_layers=layers)
    bank = ref_shim.build_reference_bank(MODALS, v, dict(C.CLIP_TEXT), projection_dim=768)
    model = ref_shim.build_reference_model(bank, 'sum', MODALS, 3, feature_dims=768, fusion_dim=256, dropout_prob=0.1,
                                           extra_missing_codes={'depth': 5, 'thermal': 6})
    sd = R.synth_state_dict([(k, tuple(t.shape)) for k, t in model.state_dict().items()])
    model.load_state_dict(sd, strict=False)
    del sd
    model.train()
    cfgs, tcfg = full_configs(layers)
    data = R.synth_inputs(MODALS, n_samples, cfgs, tcfg, seed=0)
    mi = R.synth_missing_index(n_samples, 0.3, MODALS)
    labels = torch.arange(n_samples) % 3
    crit = torch.nn.CrossEntropyLoss()

    def step():
        model.zero_grad(set_to_none=True)
        loss = crit(model({k: dict(d) for k, d in data.items()}, mi), labels)
        loss.backward()
        return float(loss.detach())
    return step


def cpu_step_fn(n_samples, layers=24):
    """The oracle PORT of the same step (oracle/restatement.py) -- the fallback when the reference tree is absent."""
    import torch
    import restatement as R
    from missm_b200 import shapes
    cfgs, tcfg = full_configs(layers)
    named = shapes.reference_named_shapes(cfgs, tcfg, MODALS, 'sum')
    sd = R.synth_state_dict([(k, s) for k, s in named if 'language' not in k])
    sd = {k: v.requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    data = R.synth_inputs(MODALS, n_samples, cfgs, tcfg, seed=0)
    mi = R.synth_missing_index(n_samples, 0.3, MODALS)
    labels = torch.arange(n_samples) % 3
    scales = {m: 2.6592 for m in MODALS}

    def step():
        logits, _ = R.finetune_forward(sd, 'sum', MODALS, data, mi, cfgs, tcfg, scales)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        for v in sd.values():
            v.grad = None
        loss.backward()
        return float(loss.detach())
    return step


def cpu_arm(n_samples, layers):
    """-> (step_fn, kind, what): the reference itself when its tree is present, else the oracle port."""
    step = reference_step_fn(n_samples, layers)
    if step is not None:
        import ref_shim
        return step, "reference", f"unmodified reference ({ref_shim.REFERENCE_ROOT}: languagebind + src.model through oracle/ref_shim.py)"
    return cpu_step_fn(n_samples, layers), "port", "oracle/restatement.py"


def run_reference_arm(a):
    """The reference's own CPU implementation of the path, all host threads, fp32, full batch through every tower
    (the reference never skips missing samples).  A step is a bounded sample of the B = 64 workload: as many samples
    (2..8) as keep the whole --steps K --warmup W run within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = a.cpu_baseline_samples
    if n <= 0:      # auto: ~1.1 samples/s on the 16 host cores of the GPU box -> about 200 s for the whole run
        n = max(2, min(8, int(220 / max(1, a.steps + a.warmup))))
    step, kind, what = cpu_arm(n, a.layers)
    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    val = n * a.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n} samples per step (bounded CPU sample of the B=64 workload)",
                   "implementation": what},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{n} samples x 3 full-size towers fwd+bwd per step, {a.steps} steps; {what}"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])), mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_gpu_arm(a):
    import torch
    import torch.distributed as dist
    import restatement as R                     # only for synthetic inputs/weights + the CPU baseline leg
    from missm_b200 import ops, shapes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ops.lib()

    cfgs, tcfg = full_configs(a.layers, a.lora_r)
    model = shapes.build_finetune(cfgs, tcfg, MODALS, 'sum', 3, 768, 256, dropout_prob=0.1)
    sd = R.synth_state_dict([(k, tuple(v.shape)) for k, v in model.state_dict().items()])
    shapes.load_named(model, sd)
    del sd
    model = model.to(dev)
    for n, p in model.named_parameters():       # the text tower is registered but unused by this config
        if 'language' in n:
            p.requires_grad_(False)
    model.train()
    net = model
    if world > 1:
        # train_ddp.py:189 (broadcast_buffers=True, find_unused_parameters=False, default gradient copies).
        # MISSM_BENCH_BUCKET_VIEW=1 is a measurement switch only (gradient_as_bucket_view=True: no per-parameter
        # copies into / out of the all-reduce buckets), reported in config.ddp
        bucket_view = os.environ.get("MISSM_BENCH_BUCKET_VIEW") is not None
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=True,
                                                        find_unused_parameters=False, gradient_as_bucket_view=bucket_view)

    B = a.batch
    host = R.synth_inputs(MODALS, B, cfgs, tcfg, seed=rank)
    host = {m: {'pixel_values': v['pixel_values'].pin_memory()} for m, v in host.items()}
    mi_host = R.synth_missing_index(B, a.missing, MODALS, seed=2025 + rank).pin_memory()
    labels_host = (torch.arange(B) % 3).pin_memory()
    data = {m: {'pixel_values': v['pixel_values'].to(dev)} for m, v in host.items()}
    mi, labels = mi_host.to(dev), labels_host.to(dev)
    crit = torch.nn.CrossEntropyLoss()
    n_missing = int((mi_host != 0).sum())
    present_sample_towers = len(MODALS) * B - n_missing

    opt = None
    if a.optimizer != "none":
        from missm_b200 import optim as moptim
        trainable = [p for p in model.parameters() if p.requires_grad]
        # the call of train_ddp.py:205 (lr 1e-4, weight_decay 0 are the script's defaults, :40-41)
        opt = (moptim.FusedAdam if a.optimizer == "fused" else torch.optim.Adam)(trainable, lr=1e-4, weight_decay=0)

    def step_resident():
        net.zero_grad(set_to_none=True)
        loss = crit(net(data, mi), labels)
        loss.backward()
        if opt is not None:
            opt.step()
        return loss

    def step_e2e():
        # the call a user makes, with HOST buffers: finetune_model.forward(pinned host tensors, host missing_index);
        # every tower uploads its own input on its own stream (bank.LanguageBind.forward), the loss comes back
        net.zero_grad(set_to_none=True)
        l_ = labels_host.to(dev, non_blocking=True)
        loss = crit(net(host, mi_host), l_)
        loss.backward()
        if opt is not None:
            opt.step()
        return float(loss.detach())                            # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        from missm_b200 import bank as _bank
        w0 = _bank.HOST_WAIT_S[0]
        h0 = time.perf_counter()
        for _ in range(steps):
            fn()
        # time the host needs to ISSUE a step = wall time of the loop minus the time it sat in the compaction
        # read-back waiting for the GPU to drain the previous step
        host_ms[0] = (time.perf_counter() - h0 - (_bank.HOST_WAIT_S[0] - w0)) * 1e3 / steps
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # max over ranks, device-timed
        return float(ms)

    for _ in range(max(a.warmup, 3)):
        step_resident()
    from missm_b200 import _lib
    L = ops.lib()
    sampler = ClockSampler(local) if rank == 0 else None
    L.missm_launch_count(1)                    # the library counts its own kernel launches
    calls0 = _lib.CALLS[0]
    ms = timed(step_resident, a.steps)
    launches = int(L.missm_launch_count(0))
    binding_calls = (_lib.CALLS[0] - calls0) / a.steps
    host_issue_ms = host_ms[0]
    clocks = sampler.stop() if sampler else None
    samples_per_s = world * B * a.steps / (ms / 1e3)

    # roofline of the dominant kernel (tcgen05 GEMM): CUDA events around every GEMM launch of one
    # more step on the launching stream (kept out of the headline so the events cost nothing there)
    import ctypes
    streams_on = model.encoder.tower_streams
    model.encoder.tower_streams = False       # one stream: every GEMM is timed alone, not overlapped with another tower's
    L.missm_gemm_profile(1)
    step_resident()
    torch.cuda.synchronize()
    L.missm_gemm_profile(0)
    model.encoder.tower_streams = streams_on
    g_ms_c, g_flop_c, n_gemm_c = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    L.missm_gemm_profile_read(ctypes.byref(g_ms_c), ctypes.byref(g_flop_c), ctypes.byref(n_gemm_c))
    g_ms, g_flop, n_gemm = g_ms_c.value, g_flop_c.value, n_gemm_c.value
    pk, pk_src = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    ach_tf = g_flop / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("gemm_dram_bytes_per_launch")

    e2e = None
    if not a.no_e2e:
        for _ in range(2):
            step_e2e()
        e_ms = timed(step_e2e, a.steps)
        h2d = sum(v['pixel_values'].numel() * 4 for v in host.values()) + mi_host.numel() * 8 + labels_host.numel() * 8
        e2e = {"value": world * B * a.steps / (e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": e_ms / a.steps}

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        del net
        torch.cuda.empty_cache()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n = a.cpu_baseline_samples if a.cpu_baseline_samples > 0 else 2
        stepc, kind, what = cpu_arm(n, a.layers)
        stepc()
        t0 = time.perf_counter()
        stepc()
        dt = time.perf_counter() - t0
        cpu_base = {"value": n / dt, "unit": "samples/s", "cores": cores, "kind": kind,
                    "sample": f"{n} samples x 3 full-size towers fwd+bwd, fp32, {what}, 1 warm-up + 1 timed"}

    if rank == 0:
        # fwd+bwd = 3 x forward flops; with a frozen (LoRA) encoder the weight gradients are not computed: 2 x
        flop_factor = 2.0 if a.lora_r else 3.0
        algo_tf = present_sample_towers * flop_factor * 162.0e9 * world * a.steps / (ms / 1e3) / 1e12
        variant = []
        if a.lora_r:
            variant.append(f"LoRA r={a.lora_r} on q/k/v/out_proj, encoder frozen (dgrad-only backward, flops counted as 2 x forward)")
        if opt is not None:
            variant.append(f"optimizer.step() inside the timed step: {type(opt).__module__}.{type(opt).__name__}, "
                           f"{sum(p.numel() for p in model.parameters() if p.requires_grad) / 1e6:.1f} M trainable parameters")
        line = {
            "metric": METRIC, "value": samples_per_s, "unit": "samples/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "missing_ratio": a.missing,
                       "missing_samples": n_missing, "fusion": "sum", "layers": a.layers, "tower_streams": bool(model.encoder.tower_streams),
                       "ddp": (None if world == 1 else "DistributedDataParallel as train_ddp.py:189" +
                               (" + gradient_as_bucket_view (measurement switch)" if os.environ.get("MISSM_BENCH_BUCKET_VIEW") else "")),
                       "host_issue_ms_per_step": host_issue_ms, "binding_calls_per_step": binding_calls,
                       "variant": "; ".join(variant) if variant else None,
                       "step": "zero_grad + forward + CrossEntropy + backward (DDP allreduce at N>1); " +
                               ("optimizer excluded (metric is fwd+bwd)" if opt is None else "optimizer.step() included"),
                       "l2": "working set >> 126 MB L2 every step (1.8 GB bf16 weights + >30 GB activations)",
                       "encoder_tflops_algorithmic": algo_tf,
                       "encoder_frac_of_bf16_peak": algo_tf / world / pk["bf16_tflops"], "peaks": pk_src},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": ach_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": ach_tf / peak_tf if peak_tf else None, "traffic": traffic,
                         "launches_per_step": n_gemm, "gemm_ms_per_step": g_ms,
                         "gemm_share_of_step": g_ms / (ms / a.steps)},
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)
