import torch, time
x = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
y = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
big = torch.empty(1024 * 1024 * 1024, dtype=torch.uint8, device='cuda')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        big.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts)//2]
print('memset 256MiB (ms): min, median', t(lambda: x.zero_()))
print('fill_(0xff) 256MiB', t(lambda: x.fill_(255)))
print('copy 256MiB->256MiB', t(lambda: y.copy_(x)))
x32 = x.view(torch.int32)
print('memset 128MiB x2', t(lambda: (x[:128*1024*1024].zero_(), y[:128*1024*1024].zero_())))
