"""Do two INT8-route products on two handles (two streams) overlap?  Wall time of two threads, each running R products on
its own handle, against one thread running 2R (development aid)."""
import os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_emu_uqsa_b200 import _lib

n, batch, R = 2048, 8, 12
devs = [_lib.Device(0), _lib.Device(0)]
g = torch.Generator(device="cuda").manual_seed(1)
data = []
for _ in devs:
    X = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g)
    Y = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g)
    C = torch.zeros(batch, n, n, dtype=torch.float64, device="cuda")
    data.append((X, Y, C))
sz = n * n


def run(i, reps, fn):
    X, Y, C = data[i]
    for _ in range(reps):
        fn(devs[i], X, Y, C)


def oz(dev, X, Y, C):
    dev.dbg_gemm_oz(X, Y, C, n, n, n, n, n, n, sz, sz, sz, batch=batch, layout=1, nmod=16)


def dm(dev, X, Y, C):
    dev.dbg_gemm(X, Y, C, n, n, n, n, n, n, sz, sz, sz, batch=batch, layout=1)


for name, fn in (("int8 route", oz), ("dmma", dm)):
    run(0, 2, fn); run(1, 2, fn)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); run(0, 2 * R, fn); torch.cuda.synchronize(); t1 = time.perf_counter() - t0
    th = [threading.Thread(target=run, args=(i, R, fn)) for i in range(2)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize(); t2 = time.perf_counter() - t0
    print(f"{name}: one stream {t1 * 1e3 / (2 * R):.3f} ms per product, two streams {t2 * 1e3 / (2 * R):.3f} ms per product")
