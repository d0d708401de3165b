#!/usr/bin/env python
"""Is GPUDirect Storage usable on this box?  Reports the nvidia-fs kernel module and what cuFileDriverOpen returns
(round 1: module absent, driver opens in compatibility mode -> DESIGN.md section 9 item 3)."""
import ctypes, os, subprocess
print("nvidia_fs module:", subprocess.run("lsmod | grep -i nvidia_fs; ls /proc/driver/nvidia-fs 2>&1 | head -2", shell=True, capture_output=True, text=True).stdout.strip() or "absent")
class Err(ctypes.Structure):
    _fields_ = [("err", ctypes.c_int), ("cu_err", ctypes.c_int)]
try:
    lib = ctypes.CDLL("/usr/local/cuda/lib64/libcufile.so")
    lib.cuFileDriverOpen.restype = Err
    import torch; torch.cuda.init(); torch.zeros(1, device="cuda")
    r = lib.cuFileDriverOpen()
    print("cuFileDriverOpen ->", r.err, r.cu_err)
    if r.err == 0:
        lib.cuFileDriverClose.restype = Err
        lib.cuFileDriverClose()
except Exception as e:
    print("cufile probe failed:", e)
