"""INT8 tensor-core GEMM rate of this B200 through the library path (torch._int_mm -> cuBLASLt), as the roof for the
INT8-slice FP64 emulation discussed in DESIGN.md section 11 (a future kernel; nothing in the product calls this).
Prints one JSON line: TOP/s for a few square sizes, and the FP64-equivalent rate for 8 / 9 / 10 slices."""
import json
import torch

assert torch.cuda.is_available()
res = {}
for n in (4096, 8192, 16384):
    a = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda")
    for _ in range(3):
        c = torch._int_mm(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        c = torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res["int8_n%d_TOPs" % n] = round(2.0 * n ** 3 / ms * 1e-9, 1)
best = max(res.values())
for s in (8, 9, 10):
    res["fp64_equiv_TFLOPs_s%d" % s] = round(best / (s * (s + 1) / 2), 1)
res["gpu"] = torch.cuda.get_device_name(0)
print(json.dumps(res))
