import torch, time
x = torch.empty(1 << 29, dtype=torch.int16, device="cuda")   # 1 GiB
y = torch.empty(1 << 29, dtype=torch.int16, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(n): fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n
t = timeit(lambda: x.fill_(7)); print("fill_ 1 GiB: %.3f ms  %.0f GB/s written" % (t, 1.0737e9 / t / 1e6))
t = timeit(lambda: x.zero_()); print("zero_ 1 GiB: %.3f ms  %.0f GB/s written" % (t, 1.0737e9 / t / 1e6))
t = timeit(lambda: y.copy_(x)); print("copy 1 GiB: %.3f ms  %.0f GB/s read+write" % (t, 2 * 1.0737e9 / t / 1e6))
s = x.sum  # read-only
t = timeit(lambda: torch.sum(x.view(torch.int32)[: 1 << 28], dtype=torch.int64)); print("sum (read 1 GiB): %.3f ms  %.0f GB/s read" % (t, 1.0737e9 / t / 1e6))
