"""int8 GEMM peak measured the way MEASURED_PEAKS.json measures bf16: a library GEMM (cuBLASLt through
torch._int_mm, 8192^3), best of 10 (burst) and back to back for 4 s (sustained).  Denominator only."""
import json, time, torch
n = 8192
a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda").t().contiguous().t()
for _ in range(3): torch._int_mm(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); reps = 0; e0.record()
while time.time() - t0 < 4.0:
    for _ in range(20): torch._int_mm(a, b)
    reps += 20; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / reps
x = torch.randn(n, n, dtype=torch.bfloat16, device="cuda"); y = torch.randn(n, n, dtype=torch.bfloat16, device="cuda")
for _ in range(3): x @ y
bb = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); x @ y; e1.record(); torch.cuda.synchronize(); bb = min(bb, e0.elapsed_time(e1))
print(json.dumps({"int8_tops_burst": 2 * n**3 / best / 1e9, "int8_tops_sustained": 2 * n**3 / sus / 1e9,
                  "bf16_tflops_burst_here": 2 * n**3 / bb / 1e9, "how": "torch._int_mm / torch.matmul 8192^3, CUDA events"}))
