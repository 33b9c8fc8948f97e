import torch, time
dev=torch.device("cuda:0")
n=1024*16385
h=torch.empty(n,dtype=torch.int64).pin_memory(); h.random_(0,1000000)
d=torch.empty(n,dtype=torch.int64,device=dev)
def t(fn,it=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
ms=t(lambda: d.copy_(h,non_blocking=True)); print("single stream 134MB: %.3f ms %.1f GB/s"%(ms, n*8/ms/1e6))
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def two():
    cur=torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): d[:n//2].copy_(h[:n//2],non_blocking=True)
    with torch.cuda.stream(s2): d[n//2:].copy_(h[n//2:],non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
ms=t(two); print("two streams: %.3f ms %.1f GB/s"%(ms, n*8/ms/1e6))
h32=torch.empty(n,dtype=torch.int32).pin_memory(); d32=torch.empty(n,dtype=torch.int32,device=dev)
ms=t(lambda: d32.copy_(h32,non_blocking=True)); print("int32 67MB: %.3f ms %.1f GB/s"%(ms, n*4/ms/1e6))
# concurrent with a bandwidth-heavy kernel
a=torch.empty(1<<28,dtype=torch.float32,device=dev); b=torch.empty_like(a)
def both():
    cur=torch.cuda.current_stream(); s1.wait_stream(cur)
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    b.copy_(a); b.copy_(a)
    cur.wait_stream(s1)
ms=t(both); print("copy under 2x device copy (%.3f ms total)"%ms)
ms2=t(lambda:(b.copy_(a),b.copy_(a))); print("2x device copy alone %.3f ms"%ms2)
