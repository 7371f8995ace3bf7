"""Host<->device copy rates of this box for the sizes the drop-in API moves (pageable vs pinned, and a host memcpy)."""
import time, numpy as np, torch
def t(fn, n=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
for mb in (0.5, 1.5, 6.6, 50, 150):
    nbytes = int(mb * 1e6); n = nbytes // 4
    hp = torch.empty(n, dtype=torch.float32); hq = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device='cuda')
    a = np.empty(n, np.float32); b = np.empty(n, np.float32)
    r = {
        'h2d_pageable': t(lambda: d.copy_(hp)), 'h2d_pinned': t(lambda: d.copy_(hq, non_blocking=True)),
        'd2h_pageable': t(lambda: hp.copy_(d)), 'd2h_pinned': t(lambda: hq.copy_(d, non_blocking=True)),
        'host_memcpy': t(lambda: np.copyto(b, a)),
    }
    print(f'{mb:6.1f} MB: ' + '  '.join(f'{k} {v*1e6:8.1f} us ({nbytes/v/1e9:5.1f} GB/s)' for k, v in r.items()))
t0 = time.perf_counter(); x = torch.empty(int(6.6e6)//4, dtype=torch.float32).pin_memory(); print('pin 6.6MB alloc', (time.perf_counter()-t0)*1e6, 'us')
