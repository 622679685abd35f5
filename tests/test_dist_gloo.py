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
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
