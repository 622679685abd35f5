"""Device timeline of one training step of the bench workload under DDP (torchrun, N ranks), taken with
torch.profiler (CUPTI kernel records only) on rank 0 -- the replacement for the nsys timeline this image lacks.
Writes gpurun_out/<tag>_timeline.json.gz (chrome trace) and prints a summary: the span of the step, busy time of
the compute kernels per stream, the NCCL kernels, and how much of the all-reduce is exposed after the last compute
kernel of the backward."""
import argparse
import gzip
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist
import restatement as R
from missm_b200 import shapes, config as C


def summarise(path, out=sys.stdout):
    op = gzip.open if path.endswith(".gz") else open
    ev = json.load(op(path, "rt"))["traceEvents"]
    ks = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ks.sort(key=lambda e: e["ts"])
    if not ks:
        print("no kernel records", file=out)
        return
    t0 = ks[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in ks)
    print(f"span {(t1 - t0) / 1e3:.2f} ms, {len(ks)} device records", file=out)

    def is_nccl(e):
        return "nccl" in e["name"].lower()

    def union(iv):
        iv = sorted(iv)
        tot, cs, ce = 0.0, None, None
        for s, e in iv:
            if cs is None:
                cs, ce = s, e
            elif s <= ce:
                ce = max(ce, e)
            else:
                tot += ce - cs
                cs, ce = s, e
        if cs is not None:
            tot += ce - cs
        return tot

    by_stream = {}
    for e in ks:
        by_stream.setdefault(e["args"].get("stream", -1), []).append(e)
    for st, es in sorted(by_stream.items(), key=lambda kv: -sum(x["dur"] for x in kv[1])):
        busy = union([(x["ts"], x["ts"] + x["dur"]) for x in es])
        n_nccl = sum(is_nccl(x) for x in es)
        first, last = es[0]["ts"] - t0, max(x["ts"] + x["dur"] for x in es) - t0
        top = {}
        for x in es:
            k = x["name"].split("<")[0].split("(")[0][-48:]
            top[k] = top.get(k, 0.0) + x["dur"]
        tops = ", ".join(f"{k} {v / 1e3:.1f}" for k, v in sorted(top.items(), key=lambda kv: -kv[1])[:3])
        print(f"stream {st}: {len(es)} records ({n_nccl} nccl), busy {busy / 1e3:.2f} ms, first {first / 1e3:.2f} "
              f"last {last / 1e3:.2f} ms | {tops}", file=out)
    comp = [e for e in ks if not is_nccl(e) and e.get("cat") == "kernel" and
            not any(s in e["name"] for s in ("elementwise", "copy", "Memcpy", "fill", "foreach"))]
    nc = [e for e in ks if is_nccl(e)]
    last_comp = max(e["ts"] + e["dur"] for e in comp)
    print(f"hand-written / compute kernels: union busy {union([(e['ts'], e['ts'] + e['dur']) for e in comp]) / 1e3:.2f} ms, "
          f"last ends at {(last_comp - t0) / 1e3:.2f} ms", file=out)
    if nc:
        nb = union([(e["ts"], e["ts"] + e["dur"]) for e in nc])
        first_n = nc[0]["ts"] - t0
        last_n = max(e["ts"] + e["dur"] for e in nc) - t0
        after = union([(max(e["ts"], last_comp), e["ts"] + e["dur"]) for e in nc if e["ts"] + e["dur"] > last_comp])
        print(f"nccl kernels: {len(nc)}, union busy {nb / 1e3:.2f} ms, first starts {first_n / 1e3:.2f}, last ends "
              f"{last_n / 1e3:.2f} ms; busy AFTER the last compute kernel {after / 1e3:.2f} ms", file=out)
        # when did each all-reduce run, in tenths of the span
        hist = [0.0] * 10
        for e in nc:
            b = min(9, int(10 * (e["ts"] - t0) / (t1 - t0)))
            hist[b] += e["dur"] / 1e3
        print("nccl busy ms per tenth of the span: " + " ".join(f"{h:.1f}" for h in hist), file=out)
    tail = [e for e in ks if e["ts"] >= last_comp]
    tt = {}
    for e in tail:
        k = e["name"].split("<")[0][-60:]
        tt[k] = tt.get(k, 0.0) + e["dur"]
    print("records that start after the last compute kernel: " +
          ", ".join(f"{k} {v / 1e3:.2f} ms" for k, v in sorted(tt.items(), key=lambda kv: -kv[1])[:6]), file=out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tag", default="ddp")
    ap.add_argument("--summarise", default=None)
    a = ap.parse_args()
    if a.summarise:
        return summarise(a.summarise)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    MODALS = ["image", "depth", "thermal"]
    cfgs = {}
    for m in MODALS:
        v = {k: val for k, val in C.VIT_L14.items() if k != 'lora_r'}
        v['num_hidden_layers'] = a.layers
        v.update(C.SYNTHETIC_PER_MODALITY[m])
        cfgs[m] = R.vision_config(**v)
    tcfg = R.text_config(**dict(C.CLIP_TEXT, num_hidden_layers=1))
    model = shapes.build_finetune(cfgs, tcfg, MODALS, 'sum', 3, 768, 256, dropout_prob=0.1)
    sd = R.synth_state_dict([(k, tuple(t.shape)) for k, t in model.state_dict().items()])
    shapes.load_named(model, sd)
    model = model.to(dev).train()
    for n, p in model.named_parameters():
        if 'language' in n:
            p.requires_grad_(False)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(
            model, device_ids=[local], broadcast_buffers=True, find_unused_parameters=False,
            gradient_as_bucket_view=os.environ.get("MISSM_DDP_BUCKET_VIEW", "1") == "1")
    B = a.batch
    data = {m: {'pixel_values': x['pixel_values'].to(dev)} for m, x in R.synth_inputs(MODALS, B, cfgs, tcfg, seed=rank).items()}
    mi = R.synth_missing_index(B, 0.3, MODALS, seed=2025 + rank).to(dev)
    labels = (torch.arange(B, device=dev) % 3)

    def step():
        net.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(net(data, mi), labels)
        loss.backward()
        return loss

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        raw = os.path.join(ROOT, "gpurun_out", f"{a.tag}_timeline.json")
        prof.export_chrome_trace(raw)
        with open(raw, "rb") as f, gzip.open(raw + ".gz", "wb") as g:
            g.write(f.read())
        os.remove(raw)
        with open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_timeline_summary.txt"), "w") as f:
            summarise(raw + ".gz", f)
        print(open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_timeline_summary.txt")).read())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
