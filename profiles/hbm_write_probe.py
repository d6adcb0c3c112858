import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
N=4<<30
x=torch.empty(N,dtype=torch.uint8,device='cuda'); y=torch.empty(N,dtype=torch.uint8,device='cuda')
ms=t(lambda: x.fill_(3)); print("fill  (write only) %.1f GB/s" % (N/ms/1e6))
ms=t(lambda: y.copy_(x)); print("copy  (r+w)        %.1f GB/s" % (2*N/ms/1e6))
xf=x.view(torch.float32)
ms=t(lambda: xf.sum()); print("sum   (read only)  %.1f GB/s" % (N/ms/1e6))
