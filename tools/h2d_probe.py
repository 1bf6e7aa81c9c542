import torch, time
n = 2123366400
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for chunk in (n, n//8, n//32):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record()
        for o in range(0, n - chunk + 1, chunk):
            d[o:o+chunk].copy_(h[o:o+chunk], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
    print('chunk %d MB: %.1f GB/s' % (chunk >> 20, n / e0.elapsed_time(e1) / 1e6))
# two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
half = n // 2
with torch.cuda.stream(s1):
    d[:half].copy_(h[:half], non_blocking=True)
with torch.cuda.stream(s2):
    d[half:].copy_(h[half:], non_blocking=True)
torch.cuda.synchronize()
print('two streams: %.1f GB/s' % (n / (time.perf_counter() - t0) / 1e9))
