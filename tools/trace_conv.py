"""Per-tile clock64 timeline of CTA 0 of the row-streaming conv (prg_test_conv_f16 with PRG_CONV_TRACE=1).
The trace stamps are compiled in only when the library was built with
    PRG_BUILD_DEFINES="-DPRG_CONV_TRACE_BUILD" python -m pointreggpt_b200.build --force
(the production kernels carry neither the tests nor the registers of the instrumentation)."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from pointreggpt_b200 import _ffi
B,H,W,Cin,Cout = 32,256,256,64,64
x = torch.randn(B,H,W,Cin, device="cuda").half(); w = (torch.randn(Cout, 9*Cin, device="cuda")*0.05).half()
bias = torch.randn(Cout, device="cuda"); y = torch.empty(B,H,W,Cout, device="cuda", dtype=torch.float16)
os.environ.pop("PRG_CONV_TRACE", None)
for _ in range(2):
    _ffi.check(_ffi.lib().prg_test_conv_f16(_ffi.ptr(x), _ffi.ptr(w), _ffi.ptr(bias), _ffi.ptr(y), B,H,W,Cin,Cout,1,_ffi.stream()))
torch.cuda.synchronize()
os.environ["PRG_CONV_TRACE"] = "1"
_ffi.check(_ffi.lib().prg_test_conv_f16(_ffi.ptr(x), _ffi.ptr(w), _ffi.ptr(bias), _ffi.ptr(y), B,H,W,Cin,Cout,1,_ffi.stream()))
torch.cuda.synchronize()
