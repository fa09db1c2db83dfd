import torch, time
dev=torch.device('cuda:0')
for nbytes in (1<<20, 2<<20, 8<<20, 64<<20):
    h=torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d=torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    t=e0.elapsed_time(e1)/20
    e0.record()
    for _ in range(20): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    t2=e0.elapsed_time(e1)/20
    print(nbytes>>20,'MB H2D',round(t*1e3,1),'us',round(nbytes/t/1e6,1),'GB/s  D2H',round(t2*1e3,1),'us',round(nbytes/t2/1e6,1),'GB/s')
t0=time.perf_counter()
for _ in range(1000): torch.cuda.current_stream().synchronize()
print('sync overhead us', (time.perf_counter()-t0)*1e3)
