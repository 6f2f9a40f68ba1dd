import torch, time
N=4000000
dev=torch.device('cuda')
x=torch.randn(2,N,dtype=torch.float64,device=dev)
for pin in (True,):
    h=torch.empty(2,N,dtype=torch.float64,pin_memory=pin)
    for i in range(3):
        torch.cuda.synchronize(); t0=time.perf_counter(); h.copy_(x,non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
        print('D2H 64MB pinned', (t1-t0)*1e3,'ms', 64e6/(t1-t0)/1e9,'GB/s')
    t0=time.perf_counter(); h2=torch.empty(2,N,dtype=torch.float64,pin_memory=True); t1=time.perf_counter(); print('alloc pinned new', (t1-t0)*1e3)
    del h2
    t0=time.perf_counter(); h2=torch.empty(2,N,dtype=torch.float64,pin_memory=True); t1=time.perf_counter(); print('alloc pinned cached', (t1-t0)*1e3)
a=torch.randn(N*3,dtype=torch.float64).numpy()
t0=time.perf_counter(); g=torch.from_numpy(a).to(dev); torch.cuda.synchronize(); t1=time.perf_counter(); print('H2D 96MB pageable',(t1-t0)*1e3,'ms')
