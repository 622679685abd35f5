"""world_size-2 `gloo` tests on CPU for the N > 1 host logic: max-over-ranks timing, per-rank
shards, DDP over a fusion head (every parameter gets a reduced gradient), and the reference arm
of bench.py under torchrun (rank 0 alone prints, the other rank exits 0)."""
import json
import os
import socket
import subprocess
import sys
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      MISSM_DDP_SMS="132")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from missm_b200 import dist_utils
    import restatement as R
    from src.model import baseline as B
    out = {}
    out["max"] = dist_utils.max_over_ranks(10.0 + rank)
    mi = R.synth_missing_index(64, 0.3, ['image', 'depth', 'thermal'], seed=dist_utils.rank_seed(2025, rank))
    out["missing"] = mi.tolist()
    # optional data-parallel policy of the bank (MISSM_DDP_SMS): the backward pass leaves SMs to NCCL, and the
    # all-reduce gets no more channels than that (set at package import: WORLD_SIZE > 1 is in the environment)
    from missm_b200.bank import _ddp_backward_sms
    out["ddp_sms"] = _ddp_backward_sms()
    out["nccl_channels"] = os.environ.get("NCCL_MAX_NCHANNELS")
    # DDP over a (pure torch) fusion head on CPU: grads must be identical on both ranks afterwards
    torch.manual_seed(0)
    args = types.SimpleNamespace(modality_types=['image', 'audio'], feature_dims=16, fusion_dim=8, dropout_prob=0.0)
    head = B.modal_concat(args, 3)
    ddp = torch.nn.parallel.DistributedDataParallel(head)
    g = torch.Generator().manual_seed(100 + rank)
    batch = {'image': torch.randn(4, 16, generator=g), 'audio': torch.randn(4, 16, generator=g)}
    loss = ddp(batch, torch.tensor([0, 4, 3, 0])).square().mean()
    loss.backward()
    out["grads"] = {n: p.grad.flatten().tolist() for n, p in head.named_parameters()}
    out["all_have_grad"] = all(p.grad is not None for p in head.parameters())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0]["max"] == res[1]["max"] == 11.0
    assert res[0]["missing"] != res[1]["missing"] and sum(1 for v in res[0]["missing"] if v) == 19
    assert res[0]["all_have_grad"] and res[1]["all_have_grad"]
    assert res[0]["ddp_sms"] == res[1]["ddp_sms"] == 132 and res[0]["nccl_channels"] == "16"
    for n in res[0]["grads"]:
        assert torch.allclose(torch.tensor(res[0]["grads"][n]), torch.tensor(res[1]["grads"][n])), n


def _worker_model(rank, world, port, q, lora_r):
    """The PRODUCT model (tiny towers + `sum` head, optionally LoRA-wrapped) under DDP exactly as train_ddp.py:189
    wraps it, over torch stand-ins of the C ABI (tests/ops_emulation.py): rank 1's batch has NO sample with an image,
    so its image tower runs zero rows and must still hand DDP a gradient for every trainable parameter."""
    sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import restatement as R
    import ops_emulation as E
    from missm_b200 import shapes
    v = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2, patch_size=14,
             image_size=28, lora_r=lora_r, lora_alpha=16)
    cfgs = {'image': R.vision_config(**v), 'depth': R.vision_config(**v)}
    tcfg = R.text_config(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2, vocab_size=50)
    modal = ['image', 'depth']
    model = shapes.build_finetune(cfgs, tcfg, modal, 'sum', 3, 32, 16)
    shapes.load_named(model, R.synth_state_dict([(k, tuple(t.shape)) for k, t in model.state_dict().items()]))
    for n, p in model.named_parameters():
        if 'language' in n:
            p.requires_grad_(False)                    # registered (text tower of the last model) but unused here
    B = 4
    data = R.synth_inputs(modal, B, cfgs, tcfg, seed=rank)
    mi = torch.tensor([0, 5, 0, 4]) if rank == 0 else torch.tensor([4, 4, 4, 4])   # rank 1: image missing everywhere
    labels = torch.arange(B) % 3
    out = {}
    with E.emulated_fp32_mode(precision="bf16", wide_bf16=True):
        ddp = torch.nn.parallel.DistributedDataParallel(model, broadcast_buffers=True, find_unused_parameters=False)
        loss = torch.nn.functional.cross_entropy(ddp(data, mi), labels)
        loss.backward()
    trainable = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    out["all_have_grad"] = all(p.grad is not None for _, p in trainable)
    out["frozen_have_none"] = all(p.grad is None for _, p in model.named_parameters() if not p.requires_grad)
    out["n_trainable"] = len(trainable)
    out["grads"] = {n: p.grad.flatten()[:64].tolist() for n, p in trainable}
    out["loss"] = float(loss)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _worker_ddp_switch(rank, world, port, q):
    """MISSM_DDP_BUCKET_VIEW=1: importing the drop-in `languagebind` makes the UNCHANGED DDP call of train_ddp.py:189
    default to bucket views and the larger bucket; an explicit keyword wins; gradients are those of the stock reducer."""
    sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      MISSM_DDP_BUCKET_VIEW="1", MISSM_DDP_BUCKET_MB="7")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import languagebind  # noqa: F401  (the switch is read at import)
    from torch.nn.parallel import DistributedDataParallel as DDP
    from src.model import baseline as B
    args = types.SimpleNamespace(modality_types=['image', 'audio'], feature_dims=16, fusion_dim=8, dropout_prob=0.0)
    out = {}

    def run(make_ddp):
        torch.manual_seed(0)
        head = B.modal_concat(args, 3)
        ddp = make_ddp(head)
        opt = torch.optim.SGD(head.parameters(), lr=0.1)
        g = torch.Generator().manual_seed(100 + rank)
        for _ in range(3):          # zero_grad (set_to_none) -> forward -> backward -> step, as train_ddp.py:222-254
            opt.zero_grad()
            batch = {'image': torch.randn(4, 16, generator=g), 'audio': torch.randn(4, 16, generator=g)}
            ddp(batch, torch.tensor([0, 4, 3, 0])).square().mean().backward()
            opt.step()
        return ddp, torch.cat([p.detach().flatten() for p in head.parameters()])

    ddp, w_switch = run(lambda m: DDP(m, broadcast_buffers=True, find_unused_parameters=False))      # the script's call
    out["bucket_view"] = bool(ddp.gradient_as_bucket_view)
    out["bucket_bytes"] = int(ddp.bucket_bytes_cap)
    ddp2, w_stock = run(lambda m: DDP(m, broadcast_buffers=True, find_unused_parameters=False,
                                      gradient_as_bucket_view=False, bucket_cap_mb=25))
    out["explicit_wins"] = (not ddp2.gradient_as_bucket_view) and int(ddp2.bucket_bytes_cap) == 25 * 1024 * 1024
    out["same_weights"] = bool(torch.allclose(w_switch, w_stock, rtol=0, atol=1e-7))
    out["weights"] = w_switch.tolist()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _run_two_ranks(target, *extra):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, 2, port, q) + extra) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


def test_two_rank_ddp_integration_switch():
    res = _run_two_ranks(_worker_ddp_switch)
    for r in (0, 1):
        assert res[r]["bucket_view"] and res[r]["bucket_bytes"] == 7 * 1024 * 1024
        assert res[r]["explicit_wins"] and res[r]["same_weights"]
    assert res[0]["weights"] == res[1]["weights"]


def test_two_rank_product_model_with_an_empty_tower():
    res = _run_two_ranks(_worker_model, 0)
    assert res[0]["all_have_grad"] and res[1]["all_have_grad"] and res[0]["loss"] != res[1]["loss"]
    for n in res[0]["grads"]:                          # all-reduced: identical on both ranks
        assert torch.allclose(torch.tensor(res[0]["grads"][n]), torch.tensor(res[1]["grads"][n]), rtol=1e-5, atol=1e-7), n


def test_two_rank_lora_model_frozen_encoder():
    """LoRA-wrapped towers under DDP: the frozen encoder weights are not DDP parameters and get no gradient, the
    adapters (and embeddings, projections, head) are reduced; the empty tower on rank 1 does not stall the reducer."""
    res = _run_two_ranks(_worker_model, 2)
    assert res[0]["all_have_grad"] and res[1]["all_have_grad"]
    assert res[0]["frozen_have_none"] and res[1]["frozen_have_none"]
    assert any('.lora_A.' in n for n in res[0]["grads"]) and not any('mlp.fc1' in n for n in res[0]["grads"])
    for n in res[0]["grads"]:
        assert torch.allclose(torch.tensor(res[0]["grads"][n]), torch.tensor(res[1]["grads"][n]), rtol=1e-5, atol=1e-7), n


def test_reference_arm_under_torchrun_prints_one_line():
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
           "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--layers", "1",
           "--cpu-baseline-samples", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["value"] > 0
    # the reference tree (or its staged copy oracle/_ref) is present wherever build() ran: the arm times the reference itself
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
    if os.path.isdir("/root/reference/languagebind") or os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "languagebind")):
        assert d["cpu_baseline"]["kind"] == "reference"
