"""Driver for ncu captures of the hot DMMA GEMM: LAUUM-shaped (TN, lower tiles, k >= ti*128) or
SYRK/TRMM-shaped launches through the debug entry, on torch device buffers.
    python tools/ncu_gemm.py [lauum|syrk|trmm] [n] [batch] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gp_emu_uqsa_b200 import _lib

kind = sys.argv[1] if len(sys.argv) > 1 else "lauum"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = _lib.Device(0)
torch.manual_seed(0)
A = torch.randn(batch, n, n, dtype=torch.float64, device="cuda").tril_()
B = torch.randn(batch, n, n, dtype=torch.float64, device="cuda")
Cm = torch.zeros(batch, n, n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
s = n * n
st = torch.cuda.ExternalStream(dev.stream_ptr)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(reps):
    e0.record(st)
    if kind == "lauum":      # C = A^T A, lower tiles, k >= ti*128
        dev.dbg_gemm(A, A, Cm, n, n, n, n, n, n, s, s, s, 1.0, 0, _lib.KM_GE_I, 1, batch, 2)
        fl = batch * n ** 3 / 3.0
    elif kind == "syrk":     # C -= B B^T, lower tiles
        dev.dbg_gemm(B, B, Cm, n, n, n, n, n, n, s, s, s, -1.0, 1, _lib.KM_FULL, 1, batch, 0)
        fl = batch * n ** 3
    else:                    # C = B A^T with A lower (k <= j)
        dev.dbg_gemm(B, A, Cm, n, n, n, n, n, n, s, s, s, 1.0, 0, _lib.KM_LE_J, 0, batch, 0)
        fl = batch * n ** 3
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%s n=%d batch=%d: %.3f ms, %.2f TFLOP/s (algorithmic)" % (kind, n, batch, ms, fl / ms * 1e-9), flush=True)
