"""argtypes of every exported entry point of include/missm_b200.h (one table, checked by
tests/test_abi.py against the header and the built library's symbol table)."""
import ctypes

P = ctypes.c_void_p
I = ctypes.c_int32
F = ctypes.c_float

SIGNATURES = {
    "missm_gemm_bf16": [P, P],
}
