"""Round-2 aid: time every GEMM shape of one LoRA attention block in isolation (CUDA events, warm), to see where the
~19 ms / step of the adapter GEMMs go (576 skinny launches per step, ~33 us each with their gaps).
    python scratch/lora_skinny_bench.py [M] [D] [r]        (needs a B200)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
from missm_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 14906
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
r = int(sys.argv[3]) if len(sys.argv) > 3 else 2
R3, R1 = (3 * r + 7) // 8 * 8, (r + 7) // 8 * 8
dev, bf, f32 = "cuda", torch.bfloat16, torch.float32
rnd = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(bf)
hcat, qkvcat, attncat = rnd(M, D + R3), rnd(M, 3 * D + R3), rnd(M, D + R1)
dycat, dqkvcat = rnd(M, D + R1), rnd(M, 3 * D + R3)
wf_qkv, wb_qkv, wf_o, wb_o = rnd(3 * D, D + R3), rnd(3 * D + R3, D), rnd(D, D + R1), rnd(D + R1, D)
cases = {
    "fwd  T = h A^T            [M,8]  K=D": lambda: ops.gemm(hcat[:, :D], wb_qkv[3 * D:], out=hcat[:, D:]),
    "fwd  qkv = [h|T][W|sB]^T  [M,3D] K=D+8": lambda: ops.gemm(hcat, wf_qkv, out=qkvcat[:, :3 * D]),
    "fwd  T_o = attn A_o^T     [M,8]  K=D": lambda: ops.gemm(attncat[:, :D], wb_o[D:], out=attncat[:, D:]),
    "fwd  out = [a|T_o][Wo|sB] [M,D]  K=D+8": lambda: ops.gemm(attncat, wf_o, out_dtype=bf),
    "bwd  dT_o = dY sB_o       [M,8]  K=D": lambda: ops.gemm(dycat[:, :D], wf_o[:, D:], b_mn=True, out=dycat[:, D:]),
    "bwd  d_attn = [dY|dT][W;A][M,D]  K=D+8": lambda: ops.gemm(dycat, wb_o, b_mn=True, out_dtype=bf),
    "bwd  dA_o = dT_o^T attn   [8,D]  K=M": lambda: ops.gemm(dycat[:, D:], attncat[:, :D], a_mn=True, b_mn=True, out_dtype=f32),
    "bwd  dB_o = dY^T T_o      [D,8]  K=M": lambda: ops.gemm(dycat[:, :D], attncat[:, D:], a_mn=True, b_mn=True, out_dtype=f32),
    "bwd  dT = dqkv sB         [M,8]  K=3D": lambda: ops.gemm(dqkvcat[:, :3 * D], wf_qkv[:, D:], b_mn=True, out=dqkvcat[:, 3 * D:]),
    "bwd  d_h = [dqkv|dT][W;A] [M,D]  K=3D+8": lambda: ops.gemm(dqkvcat, wb_qkv, b_mn=True, out_dtype=bf),
    "bwd  dA = dT^T h          [8,D]  K=M": lambda: ops.gemm(dqkvcat[:, 3 * D:], hcat[:, :D], a_mn=True, b_mn=True, out_dtype=f32),
    "bwd  dB = dqkv^T T        [3D,8] K=M": lambda: ops.gemm(dqkvcat[:, :3 * D], hcat[:, D:], a_mn=True, b_mn=True, out_dtype=f32),
    "ref  plain qkv            [M,3D] K=D": lambda: ops.gemm(hcat[:, :D], wb_qkv[:3 * D], out_dtype=bf),
}
for name, fn in cases.items():
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us")
