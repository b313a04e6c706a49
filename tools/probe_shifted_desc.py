"""Hardware probe: UMMA K-major SW128 A-descriptor whose start is shifted by r0 rows of 128 B inside a TMA tile."""
import ctypes
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vcd_b200

vcd_b200._lib.lib()   # libvcd_b200.so provides make_act_map / error plumbing to the probe library
_here = os.path.dirname(os.path.abspath(__file__))
_so = os.path.join(_here, "_probe.so")
if not os.path.exists(_so):   # the probe kernel is NOT part of the product library: built on demand
    import subprocess
    _pkg = os.path.dirname(vcd_b200.LIB_PATH)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler",
                           "-fPIC", os.path.join(_here, "umma_probe.cu"), "-o", _so, "-L" + _pkg, "-l:libvcd_b200.so",
                           "-Xlinker", "-rpath=" + _pkg, "-lcuda"])
lib = C.CDLL(_so)
fn = lib.vcd_debug_umma_shifted
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
torch.manual_seed(0)
A = torch.randn(144, 64, device="cuda").to(torch.bfloat16)
B = torch.randn(64, 64, device="cuda").to(torch.bfloat16)
for mode in (0, 1):
    for r0 in range(0, 17):
        out = torch.zeros(128, 64, device="cuda")
        rc = fn(A.data_ptr(), B.data_ptr(), out.data_ptr(), r0, mode, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = A[r0:r0 + 128].float() @ B.float().t()
        err = float((out - ref).abs().max() / ref.abs().max())
        print(f"base_offset_mode={mode} r0={r0:2d} rc={rc} max-rel-err={err:.3e} {'OK' if err < 1e-2 else 'WRONG'}")
