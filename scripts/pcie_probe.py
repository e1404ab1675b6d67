import torch, time
dev=torch.device("cuda:0")
for mb in (1, 4, 13.6, 64):
    n=int(mb*1e6)
    d=torch.empty(n,dtype=torch.uint8,device=dev); h=torch.empty(n,dtype=torch.uint8,pin_memory=True)
    for name,fn in (("D2H",lambda: h.copy_(d,non_blocking=True)),("H2D",lambda: d.copy_(h,non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t=time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
        print(f"{name} {mb} MB: {n/dt/1e9:.1f} GB/s  {dt*1e6:.0f} us")
# bidirectional
n=int(13.6e6); d=torch.empty(n,dtype=torch.uint8,device=dev); h=torch.empty(n,dtype=torch.uint8,pin_memory=True)
d2=torch.empty(n,dtype=torch.uint8,device=dev); h2=torch.empty(n,dtype=torch.uint8,pin_memory=True)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
print(f"bidir 13.6 MB each: {n/dt/1e9:.1f} GB/s per direction")
