import torch, time
torch.cuda.init()
n_in = 4_190_000_000; n_out = 2_724_000_000
hin = torch.empty(n_in, dtype=torch.uint8).pin_memory()
hout = torch.empty(n_out, dtype=torch.uint8).pin_memory()
din = torch.empty(n_in, dtype=torch.uint8, device="cuda")
dout = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
def both(): h2d(); d2h()
a = t(h2d); b = t(d2h); c = t(both)
print("h2d %.1f ms %.1f GB/s | d2h %.1f ms %.1f GB/s | both %.1f ms" % (a, n_in / a / 1e6, b, n_out / b / 1e6, c))
# chunked 8
def chunked():
    k = 8
    for i in range(k):
        with torch.cuda.stream(s1): din[i * n_in // k:(i + 1) * n_in // k].copy_(hin[i * n_in // k:(i + 1) * n_in // k], non_blocking=True)
        with torch.cuda.stream(s2): hout[i * n_out // k:(i + 1) * n_out // k].copy_(dout[i * n_out // k:(i + 1) * n_out // k], non_blocking=True)
print("chunked both %.1f ms" % t(chunked))
