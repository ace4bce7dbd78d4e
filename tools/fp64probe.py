import ctypes as C, sys
sys.path.insert(0,'/root/repo')
from audiosdr_b200 import api
import torch
torch.zeros(1,device='cuda')
lib=api.load_library()
lib.sdrk_fp32_peak.argtypes=[C.c_int,C.c_int,C.POINTER(C.c_double),C.POINTER(C.c_float)]
for kind,name,it in ((0,'ffma',4096),(1,'fmul+fadd',4096),(2,'dfma(+2 cvt)',128)):
    ips=C.c_double(); ms=C.c_float()
    rc=lib.sdrk_fp32_peak(kind,it,C.byref(ips),C.byref(ms))
    print(name, rc, '%.3e lane-instr/s'%ips.value, '%.3f ms'%ms.value, 'per SM per cycle: %.2f'%(ips.value/148/1.965e9))
