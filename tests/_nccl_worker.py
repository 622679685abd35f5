"""Worker of tests/test_dist_nccl_gpu.py (launched with torchrun, one rank per GPU, NCCL).

SURVEY.md section 4: "2/4/8-GPU DDP run asserts identical parameters across ranks after N steps and loss curve equal
to a 1-GPU run at the same global batch".  Every rank builds the same model (name-seeded weights), wraps it in
DistributedDataParallel exactly as train_ddp.py:189 does, takes its shard of one global batch (different missing
patterns per rank, one tower with NO present sample on rank 0) and runs `steps` SGD steps; rank 0 then replays the same
steps on an unwrapped copy over the whole global batch.  Checked: gradients identical across ranks (all-reduced),
equal to the single-process gradients at the global batch, parameters identical across ranks after the steps, loss
curves equal.  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import restatement as R  # noqa: E402  (synthetic weights / inputs only)
from missm_b200 import shapes  # noqa: E402


def build(dev):
    v = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=4, patch_size=14, image_size=56)
    modal = ['image', 'depth', 'thermal']
    cfgs = {m: R.vision_config(**v) for m in modal}
    tcfg = R.text_config(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2, vocab_size=100)
    model = shapes.build_finetune(cfgs, tcfg, modal, 'sum', 3, 128, 64)
    sd = R.synth_state_dict([(k, tuple(t.shape)) for k, t in model.state_dict().items()])
    shapes.load_named(model, sd)
    model = model.to(dev)
    for n, p in model.named_parameters():
        if 'language' in n:
            p.requires_grad_(False)
    return model.train(), modal, cfgs, tcfg


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    model, modal, cfgs, tcfg = build(dev)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=True,
                                                    find_unused_parameters=False)
    per = 6
    G = per * world
    data = R.synth_inputs(modal, G, cfgs, tcfg, seed=77)
    mi = R.synth_missing_index(G, 0.5, modal, seed=5)
    mi[:per] = torch.tensor([4, 4, 4, 4, 4, 4])          # rank 0: the image tower sees NO present sample (_ZeroTower)
    labels = torch.arange(G) % 3
    sl = slice(rank * per, (rank + 1) * per)
    mine = {m: {'pixel_values': d['pixel_values'][sl].to(dev)} for m, d in data.items()}
    steps, lr = 3, 0.05
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=lr)
    losses, first_grads = [], None
    for s in range(steps):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(ddp(mine, mi[sl].to(dev)), labels[sl].to(dev))
        loss.backward()
        if s == 0:
            first_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        opt.step()
        l = loss.detach().clone()
        dist.all_reduce(l)
        losses.append((l / world).item())
    # gradients identical across ranks
    worst_rank_diff = 0.0
    for n in sorted(first_grads):
        g = first_grads[n]
        lo, hi = g.clone(), g.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        worst_rank_diff = max(worst_rank_diff, (hi - lo).abs().max().item())
    # parameters identical across ranks after the steps
    worst_param_diff = 0.0
    for n, p in model.named_parameters():
        lo, hi = p.detach().clone(), p.detach().clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        worst_param_diff = max(worst_param_diff, (hi - lo).abs().max().item())
    out = None
    if rank == 0:
        # the same steps in ONE process over the whole global batch (mean loss over G = mean of the ranks' means)
        ref, _, _, _ = build(dev)
        ropt = torch.optim.SGD([p for p in ref.parameters() if p.requires_grad], lr=lr)
        whole = {m: {'pixel_values': d['pixel_values'].to(dev)} for m, d in data.items()}
        ref_losses, worst_grad, worst_name = [], 0.0, None
        for s in range(steps):
            ropt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(ref(whole, mi.to(dev)), labels.to(dev))
            loss.backward()
            if s == 0:
                for n, p in ref.named_parameters():
                    if p.grad is None:
                        continue
                    e = ((p.grad - first_grads[n]).norm() / p.grad.norm().clamp_min(1e-20)).item()
                    if e > worst_grad:
                        worst_grad, worst_name = e, n
            ropt.step()
            ref_losses.append(loss.item())
        out = {"world": world, "losses_ddp": losses, "losses_single": ref_losses, "worst_rank_grad_diff": worst_rank_diff,
               "worst_rank_param_diff": worst_param_diff, "worst_grad_rel_vs_single": worst_grad,
               "worst_grad_name": worst_name, "n_grads": len(first_grads)}
        print("NCCL_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
