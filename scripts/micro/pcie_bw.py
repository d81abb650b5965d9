import torch, time
dev = torch.device('cuda:0')
n = 201326592 // 4
h1 = torch.empty(n, dtype=torch.float32, pin_memory=True); h2 = torch.empty(n, dtype=torch.float32, pin_memory=True)
d1 = torch.empty(n, device=dev); d2 = torch.empty(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
gb = n * 4 / 1e9
print("H2D alone  %.1f GB/s" % (gb / timed(h2d)))
print("D2H alone  %.1f GB/s" % (gb / timed(d2h)))
t = timed(both)
print("both concurrently: %.2f ms for %.0f MB each way -> %.1f GB/s per direction, %.1f GB/s total" % (t * 1e3, gb * 1e3, gb / t, 2 * gb / t))
