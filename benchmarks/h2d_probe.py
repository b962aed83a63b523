import torch, ctypes, time, numpy as np
torch.cuda.init()
n = 64*1024*1024
dev = torch.device('cuda')
dst = torch.empty(n, dtype=torch.uint8, device=dev)
def bw(src, label):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(label, n*10/ (e0.elapsed_time(e1)*1e-3)/1e9, 'GB/s')
a = torch.empty(n, dtype=torch.uint8).pin_memory(); a.fill_(1)
bw(a, 'torch pinned')
cudart = ctypes.CDLL('libcudart.so')
for flags,label in ((0,'cudaHostAlloc default'),(4,'cudaHostAlloc WC'),(1,'portable')):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags))
    buf = (ctypes.c_uint8 * n).from_address(p.value)
    t = torch.frombuffer(buf, dtype=torch.uint8)
    t.fill_(1)
    print(label, 'rc', rc, 'is_pinned', t.is_pinned())
    bw(t, label)
import os
print('cpus', os.cpu_count()); os.system('nvidia-smi topo -m | head -8; numactl -H 2>/dev/null | head -5; lspci 2>/dev/null | grep -i nvidia | head -2')
