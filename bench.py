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
FWD_GFLOP = {'image': 162.0, 'depth': 162.0, 'thermal': 162.0, 'audio': 393.4, 'video': 1711.7, 'language': 13.3}
# BASELINE.json configs by index.  1 is the headline (and the default); 2-4 are measured with `--config N` and
# reported under profiles/ (north_star: "throughput is reported at 1, 2, 4 and 8 GPUs").
CONFIGS = {
    1: dict(modals=['image', 'depth', 'thermal'], batch=64, train=True,
            workload="configs[1]: image+depth+thermal ViT-L/14 towers + sum fusion head, fwd+bwd bf16, B=64/GPU, 30% missing"),
    2: dict(modals=['audio', 'video'], batch=32, train=True,
            workload="configs[2]: audio (112x1036 mel, 593 tokens) + 8-frame video ViT-L/14 towers + sum fusion head, "
                     "fwd+bwd bf16, B=32/GPU, 30% missing"),
    3: dict(modals=['video', 'audio', 'image', 'depth', 'thermal', 'language'], batch=16, train=True,
            workload="configs[3]: all five modality towers + text + sum fusion head, fwd+bwd bf16 under DDP "
                     "(train_ddp.py:189), B=16/GPU, 30% missing"),
    4: dict(modals=['image', 'depth', 'thermal'], batch=64, train=False,
            workload="configs[4]: test.py-style no_grad eval sweep over missing rates 0-90% (10 passes of B=64 per step) "
                     "with mask compaction, image+depth+thermal + sum head, one replica per GPU"),
}
MODALS = CONFIGS[1]['modals']
WORKLOAD = CONFIGS[1]['workload']
SWEEP = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]      # data_loader.py:348,354 / test.py:119-144


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="missm", choices=["missm", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="index into BASELINE.json configs "
                    "(1 = headline: image+depth+thermal; 2 = audio+video; 3 = all five towers + text; 4 = eval sweep)")
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (0 = the config's own)")
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


def full_configs(layers=24, lora_r=0, modals=None):
    import restatement as R
    from missm_b200 import config as C
    v = {k: val for k, val in C.VIT_L14.items() if k != 'lora_r'}
    v['num_hidden_layers'] = layers
    if lora_r:
        v['lora_r'], v['lora_alpha'] = lora_r, 16
    t = dict(C.CLIP_TEXT)
    cfgs = {}
    for m in (modals or MODALS):
        if m == 'language':
            continue
        d = dict(v)
        d.update(C.SYNTHETIC_PER_MODALITY[m])
        d['temporal_mlp'] = (m != 'video')
        cfgs[m] = R.vision_config(**d)
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
def reference_step_fn(n_samples, layers=24, modals=None, train=True):
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
    modals = modals or MODALS
    vis = [m for m in modals if m != 'language']
    v = dict(C.VIT_L14, lora_r=0, num_hidden_layers=layers)
    bank = ref_shim.build_reference_bank(vis, v, dict(C.CLIP_TEXT), projection_dim=768,
                                         per_modality_cfg={m: C.SYNTHETIC_PER_MODALITY[m] for m in vis})
    model = ref_shim.build_reference_model(bank, 'sum', modals, 3, feature_dims=768, fusion_dim=256, dropout_prob=0.1,
                                           extra_missing_codes={'depth': 5, 'thermal': 6})
    sd = R.synth_state_dict([(k, tuple(t.shape)) for k, t in model.state_dict().items()])
    model.load_state_dict(sd, strict=False)
    del sd
    model.train(train)
    cfgs, tcfg = full_configs(layers, modals=modals)
    data = R.synth_inputs(modals, n_samples, cfgs, tcfg, seed=0)
    labels = torch.arange(n_samples) % 3
    crit = torch.nn.CrossEntropyLoss()
    if not train:      # eval sweep: one no_grad pass per missing rate
        mis = [R.synth_missing_index(n_samples, r, modals) for r in SWEEP]

        def sweep():
            with torch.no_grad():
                return sum(float(model({k: dict(d) for k, d in data.items()}, mi).sum()) for mi in mis)
        return sweep
    mi = R.synth_missing_index(n_samples, 0.3, modals)

    def step():
        model.zero_grad(set_to_none=True)
        loss = crit(model({k: dict(d) for k, d in data.items()}, mi), labels)
        loss.backward()
        return float(loss.detach())
    return step


def cpu_step_fn(n_samples, layers=24, modals=None, train=True):
    """The oracle PORT of the same step (oracle/restatement.py) -- the fallback when the reference tree is absent."""
    import torch
    import restatement as R
    from missm_b200 import shapes
    modals = modals or MODALS
    cfgs, tcfg = full_configs(layers, modals=modals)
    named = shapes.reference_named_shapes(cfgs, tcfg, modals, 'sum')
    sd = R.synth_state_dict([(k, s) for k, s in named if 'language' in modals or 'language' not in k])
    sd = {k: v.requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    data = R.synth_inputs(modals, n_samples, cfgs, tcfg, seed=0)
    mi = R.synth_missing_index(n_samples, 0.3, modals)
    labels = torch.arange(n_samples) % 3
    scales = {m: 2.6592 for m in cfgs}
    if not train:
        mis = [R.synth_missing_index(n_samples, r, modals) for r in SWEEP]

        def sweep():
            with torch.no_grad():
                return sum(float(R.finetune_forward(sd, 'sum', modals, data, m_, cfgs, tcfg, scales)[0].sum()) for m_ in mis)
        return sweep

    def step():
        logits, _ = R.finetune_forward(sd, 'sum', modals, data, mi, cfgs, tcfg, scales)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        for v in sd.values():
            v.grad = None
        loss.backward()
        return float(loss.detach())
    return step


def cpu_arm(n_samples, layers, modals=None, train=True):
    """-> (step_fn, kind, what): the reference itself when its tree is present, else the oracle port."""
    step = reference_step_fn(n_samples, layers, modals, train)
    if step is not None:
        import ref_shim
        return step, "reference", f"unmodified reference ({ref_shim.REFERENCE_ROOT}: languagebind + src.model through oracle/ref_shim.py)"
    return cpu_step_fn(n_samples, layers, modals, train), "port", "oracle/restatement.py"


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
    cf = CONFIGS[a.config]
    n = a.cpu_baseline_samples
    if n <= 0:      # auto: ~1.1 samples/s on the 16 host cores of the GPU box -> about 200 s for the whole run
        n = max(2, min(8, int(220 / max(1, a.steps + a.warmup))))
        if a.config != 1:
            n = 2   # the video tower alone is 10 x an image tower
    step, kind, what = cpu_arm(n, a.layers, cf['modals'], cf['train'])
    per_step = n * (len(SWEEP) if not cf['train'] else 1)
    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    val = per_step * a.steps / dt
    line = {
        "impl": "reference", "metric": METRIC if cf['train'] else "eval samples/sec (no_grad forward, missing-rate sweep)",
        "value": val, "unit": "samples/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cf['workload'], "sample": f"{n} samples per step (bounded CPU sample of the workload's batch)",
                   "implementation": what},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{n} samples x {len(cf['modals'])} full-size towers {'fwd+bwd' if cf['train'] else 'eval sweep'} per step, {a.steps} steps; {what}"},
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
    from missm_b200 import dist_utils, ops, shapes

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

    cf = CONFIGS[a.config]
    MODALS, WORKLOAD, train = cf['modals'], cf['workload'], cf['train']
    cfgs, tcfg = full_configs(a.layers, a.lora_r, MODALS)
    model = shapes.build_finetune(cfgs, tcfg, MODALS, 'sum', 3, 768, 256, dropout_prob=0.1)
    sd = R.synth_state_dict([(k, tuple(v.shape)) for k, v in model.state_dict().items()])
    shapes.load_named(model, sd)
    del sd
    model = model.to(dev)
    if 'language' not in MODALS:
        for n, p in model.named_parameters():       # the text tower is registered but unused by this config
            if 'language' in n:
                p.requires_grad_(False)
    model.train(train)
    net = model
    if world > 1 and train:
        # The literal call of train_ddp.py:189.  The package's DDP integration switch (MISSM_DDP_BUCKET_VIEW, see
        # missm_b200/ddp_integration.py and INTEGRATION.md) is ON by default here, as a deployment would set it:
        # the unchanged call then defaults to gradient_as_bucket_view=True.  MISSM_DDP_BUCKET_VIEW=0 measures the
        # stock reducer; MISSM_BENCH_BUCKET_MB is a measurement switch (bucket_cap_mb).  Both are reported in config.ddp.
        if os.environ.get("MISSM_DDP_BUCKET_VIEW", "1") == "1":
            from missm_b200 import ddp_integration
            ddp_integration.install()
        extra = {}
        if os.environ.get("MISSM_BENCH_BUCKET_MB"):
            extra["bucket_cap_mb"] = int(os.environ["MISSM_BENCH_BUCKET_MB"])
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=True,
                                                        find_unused_parameters=False, **extra)

    B = a.batch if a.batch > 0 else cf['batch']
    host = R.synth_inputs(MODALS, B, cfgs, tcfg, seed=rank)
    host = {m: {k: t.pin_memory() for k, t in v.items()} for m, v in host.items()}
    mi_host = R.synth_missing_index(B, a.missing, MODALS, seed=dist_utils.rank_seed(2025, rank)).pin_memory()
    labels_host = (torch.arange(B) % 3).pin_memory()
    data = {m: {k: t.to(dev) for k, t in v.items()} for m, v in host.items()}
    mi, labels = mi_host.to(dev), labels_host.to(dev)
    crit = torch.nn.CrossEntropyLoss()
    n_missing = int((mi_host != 0).sum())
    code_of = {'language': 1, 'video': 2, 'audio': 3, 'image': 4, 'depth': 5, 'thermal': 6}
    # forward GFLOP of the samples the towers actually run (mask compaction skips the missing ones)
    fwd_gflop_step = sum(FWD_GFLOP[m] * int((mi_host != code_of[m]).sum()) for m in MODALS)
    sweep_host = [R.synth_missing_index(B, r, MODALS, seed=dist_utils.rank_seed(2025, rank)).pin_memory() for r in SWEEP]
    sweep_dev = [t.to(dev) for t in sweep_host]
    if not train:
        fwd_gflop_step = sum(FWD_GFLOP[m] * int((t != code_of[m]).sum()) for t in sweep_host for m in MODALS)
    samples_per_step = B * (len(SWEEP) if not train else 1)

    opt = None
    if a.optimizer != "none":
        from missm_b200 import optim as moptim
        trainable = [p for p in model.parameters() if p.requires_grad]
        # the call of train_ddp.py:205 (lr 1e-4, weight_decay 0 are the script's defaults, :40-41)
        opt = (moptim.FusedAdam if a.optimizer == "fused" else torch.optim.Adam)(trainable, lr=1e-4, weight_decay=0)

    def step_resident():
        if not train:           # test.py:119-144: one no_grad pass over the batch per missing rate
            with torch.no_grad():
                return sum(net(data, m_).sum() for m_ in sweep_dev)
        net.zero_grad(set_to_none=True)
        loss = crit(net(data, mi), labels)
        loss.backward()
        if opt is not None:
            opt.step()
        return loss

    def step_e2e():
        # the call a user makes, with HOST buffers: finetune_model.forward(pinned host tensors, host missing_index);
        # every tower uploads its own input on its own stream (bank.LanguageBind.forward), the loss comes back
        if not train:
            with torch.no_grad():
                return float(sum(net(host, m_).sum() for m_ in sweep_host))
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
        return dist_utils.max_over_ranks(e0.elapsed_time(e1), device=dev)      # device-timed, max over ranks

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
    samples_per_s = world * samples_per_step * a.steps / (ms / 1e3)

    # roofline of the dominant kernel (tcgen05 GEMM): CUDA events around every GEMM launch of one
    # more step on the launching stream (kept out of the headline so the events cost nothing there)
    import ctypes
    streams_on, lockstep_on = model.encoder.tower_streams, model.encoder.lockstep
    model.encoder.tower_streams = False       # one stream: every GEMM is timed alone, not overlapped with another tower's
    model.encoder.lockstep = False            # ... and tower by tower, as the kernels of a tower follow one another
    L.missm_gemm_profile(1)
    step_resident()
    torch.cuda.synchronize()
    L.missm_gemm_profile(0)
    model.encoder.tower_streams, model.encoder.lockstep = streams_on, lockstep_on
    g_ms_c, g_flop_c, n_gemm_c = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    L.missm_gemm_profile_read(ctypes.byref(g_ms_c), ctypes.byref(g_flop_c), ctypes.byref(n_gemm_c))
    g_ms, g_flop, n_gemm = g_ms_c.value, g_flop_c.value, n_gemm_c.value
    pk, pk_src = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    ach_tf = g_flop / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp) and a.config == 1:
        with open(tp) as f:
            traffic = json.load(f).get("gemm_dram_bytes_per_launch")

    e2e = None
    if not a.no_e2e:
        for _ in range(2):
            step_e2e()
        e_ms = timed(step_e2e, a.steps)
        h2d = sum(t.numel() * t.element_size() for v in host.values() for t in v.values())
        h2d = (h2d + mi_host.numel() * 8) * (len(SWEEP) if not train else 1) + labels_host.numel() * 8
        e2e = {"value": world * samples_per_step * a.steps / (e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": e_ms / a.steps}

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        del net
        torch.cuda.empty_cache()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n = a.cpu_baseline_samples if a.cpu_baseline_samples > 0 else 2
        # in its own process: the reference's `languagebind` / `src` packages share their names with the drop-in
        # packages this process has imported
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", str(a.config),
                            "--layers", str(a.layers), "--steps", "1", "--warmup", "1", "--cpu-baseline-samples", str(n)],
                           capture_output=True, text=True, timeout=900)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode == 0 and lines:
            cpu_base = json.loads(lines[-1])["cpu_baseline"]
            cpu_base["sample"] = cpu_base["sample"].replace("1 steps", "1 warm-up + 1 timed step")
        else:
            cpu_base = {"value": None, "unit": "samples/s", "cores": cores, "kind": "unavailable",
                        "sample": "the CPU leg failed: " + r.stderr[-300:]}

    if rank == 0:
        # fwd+bwd = 3 x forward flops; with a frozen (LoRA) encoder the weight gradients are not computed: 2 x
        flop_factor = (2.0 if a.lora_r else 3.0) if train else 1.0
        algo_tf = fwd_gflop_step * 1e9 * flop_factor * world * a.steps / (ms / 1e3) / 1e12
        variant = []
        if a.lora_r:
            variant.append(f"LoRA r={a.lora_r} on q/k/v/out_proj, encoder frozen (dgrad-only backward, flops counted as 2 x forward)")
        if opt is not None:
            variant.append(f"optimizer.step() inside the timed step: {type(opt).__module__}.{type(opt).__name__}, "
                           f"{sum(p.numel() for p in model.parameters() if p.requires_grad) / 1e6:.1f} M trainable parameters")
        line = {
            "metric": METRIC if train else "eval samples/sec (no_grad forward, missing-rate sweep)", "value": samples_per_s, "unit": "samples/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "missing_ratio": a.missing,
                       "missing_samples": n_missing, "fusion": "sum", "layers": a.layers, "tower_streams": bool(model.encoder.tower_streams),
                       "ddp": (None if world == 1 or not train else "DistributedDataParallel as train_ddp.py:189" +
                               (" + MISSM_DDP_BUCKET_VIEW=1 (package switch: gradient_as_bucket_view defaults to True, "
                                 f"bucket_cap_mb to {os.environ.get('MISSM_DDP_BUCKET_MB', '200')})"
                                if os.environ.get("MISSM_DDP_BUCKET_VIEW", "1") == "1" else " (stock reducer)") +
                               (f" + bucket_cap_mb={os.environ['MISSM_BENCH_BUCKET_MB']} (measurement switch)"
                                if os.environ.get("MISSM_BENCH_BUCKET_MB") else "")),
                       "host_issue_ms_per_step": host_issue_ms, "binding_calls_per_step": binding_calls,
                       "variant": "; ".join(variant) if variant else None,
                       "baseline_config_index": a.config,
                       "step": ("10 no_grad forward passes (missing rates 0-90%), independent replicas at N>1" if not train else
                                "zero_grad + forward + CrossEntropy + backward (DDP allreduce at N>1); " +
                                ("optimizer excluded (metric is fwd+bwd)" if opt is None else "optimizer.step() included")),
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
