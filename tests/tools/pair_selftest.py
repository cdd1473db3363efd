import ctypes, sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "st-nerf_b200"))
import torch
from stnerf_b200 import _lib as L
torch.zeros(1, device="cuda")
for i in range(3):
    e = ctypes.c_float(-1)
    rc = L.lib().stnerf_selftest_umma_pair(ctypes.byref(e))
    print("pair selftest rc", rc, "max_err", e.value, flush=True)
    if rc != 0:
        print(L.lib().stnerf_last_cuda_error()); break
e = ctypes.c_float(-1)
print("single", L.lib().stnerf_selftest_umma(ctypes.byref(e)), e.value)
