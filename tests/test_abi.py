"""CPU-only: the C-ABI library loads and exports every symbol include/missm_b200.h declares, and
the ctypes table (missm_b200/_abi.py) matches the header's argument counts.  No compute calls."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "missm_b200.h")
LIB = os.path.join(ROOT, "missm-benchmark_b200", "lib", "libmissm_b200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(missm_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    return LIB


def test_header_declares_the_expected_surface():
    fns = declared_functions()
    for name in ("missm_gemm_bf16", "missm_attention_fwd", "missm_attention_bwd", "missm_layernorm_fwd",
                 "missm_layernorm_bwd", "missm_compact_mask", "missm_scatter_rows_zero", "missm_gather_rows",
                 "missm_patchify", "missm_fusion_sum_fwd", "missm_fusion_sum_bwd", "missm_version",
                 "missm_last_error"):
        assert name in fns, name


def test_library_exports_every_declared_symbol(built):
    fns = declared_functions()
    syms = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT\s+(missm_\w+)", syms))
    missing = sorted(set(fns) - exported)
    assert not missing, f"declared in the header but not exported: {missing}"
    undeclared = sorted(exported - set(fns))
    assert not undeclared, f"exported but not declared in the header: {undeclared}"


def test_ctypes_table_matches_header(built):
    from missm_b200 import _abi
    fns = declared_functions()
    table = set(_abi.SIGNATURES) | set(_abi.OTHER_EXPORTS)
    assert table == set(fns), (sorted(table - set(fns)), sorted(set(fns) - table))
    for name, argtypes in _abi.SIGNATURES.items():
        assert len(argtypes) == fns[name], (name, len(argtypes), fns[name])


def test_library_loads_and_reports_version(built):
    from missm_b200 import _lib
    L = _lib.lib()
    from missm_b200 import _abi
    hdr = int(re.search(r"#define\s+MISSM_ABI_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    assert L.missm_version() == hdr == _abi.ABI_VERSION
    assert isinstance(L.missm_last_error(), bytes)


def test_struct_mirrors_have_the_header_sizes(built):
    """sizeof of the ctypes mirrors == sizeof computed by the C compiler from the header."""
    from missm_b200._lib import AttnArgs, GemmArgs
    from missm_b200.fusion_ops import FusionSumArgs
    from missm_b200.optim import AdamArgs
    from missm_b200.ops import PreprocArgs
    from missm_b200.blocks import AttnBlockArgs, MlpBlockArgs
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "missm_b200.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(missm_gemm_args), sizeof(missm_attn_args), sizeof(missm_fusion_sum_args), sizeof(missm_adam_args), sizeof(missm_preproc_args), sizeof(missm_attn_block_args), sizeof(missm_mlp_block_args), offsetof(missm_attn_block_args, x), offsetof(missm_attn_block_args, scratch), offsetof(missm_mlp_block_args, x), offsetof(missm_mlp_block_args, scratch)); return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(GemmArgs), ctypes.sizeof(AttnArgs), ctypes.sizeof(FusionSumArgs),
                     ctypes.sizeof(AdamArgs), ctypes.sizeof(PreprocArgs), ctypes.sizeof(AttnBlockArgs),
                     ctypes.sizeof(MlpBlockArgs), AttnBlockArgs.x.offset, AttnBlockArgs.scratch.offset,
                     MlpBlockArgs.x.offset, MlpBlockArgs.scratch.offset]
