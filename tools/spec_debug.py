import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from audiosdr_b200 import aux
from oracle import aux_lib as A
dev=torch.device('cuda:0')
for nch in (70, 4096, 65536):
    g=torch.Generator(device=dev); g.manual_seed(99)
    I=(torch.randn((nch,256),generator=g,device=dev)*3000).round().to(torch.int16); Q=(torch.randn((nch,256),generator=g,device=dev)*3000).round().to(torch.int16)
    t=torch.arange(256,device=dev,dtype=torch.float32)
    I+=(8000*torch.cos(2*np.pi*37/256*t)).round().to(torch.int16)[None,:]; Q+=(8000*torch.sin(2*np.pi*37/256*t)).round().to(torch.int16)[None,:]
    h=aux.GrabberBatch(nch); h.process(I,Q,n_blocks=2)
    p=torch.zeros((nch,256),dtype=torch.float32,device=dev)
    ok=h.spectrum_device(p); torch.cuda.synchronize()
    pick=[0,1,nch-1]
    snap=h.grab(pick); want=A.grab_spectrum(snap); got=p[pick].cpu().numpy()
    hs=h.spectrum(pick)
    print(nch, ok, 'dev==oracle', np.array_equal(got,want), 'host==oracle', np.array_equal(hs,want), 'argmax', got[0].argmax(), want[0].argmax(), got[0][:3], want[0][:3])
    h.close()
