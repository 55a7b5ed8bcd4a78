# cuBLAS DGEMM peak via torch.matmul (fp64) -- the FP64 tensor roofline denominator.
import torch, json, time
dev = "cuda:0"
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2 * n**3 / best * 1e-9
    # sustained
    t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); k = 0
    while k < 10: c = a @ b; k += 1
    e1.record(); torch.cuda.synchronize()
    res[f"dgemm_{n}_tflops_sustained10"] = 10 * 2 * n**3 / e0.elapsed_time(e1) * 1e-9
# cholesky + triangular solve via cuSOLVER for context
for n in (2048, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); a = a @ a.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    for _ in range(2): l = torch.linalg.cholesky(a)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); l = torch.linalg.cholesky(a); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res[f"cusolver_potrf_{n}_ms"] = ms; res[f"cusolver_potrf_{n}_tflops"] = n**3 / 3 / ms * 1e-9
print(json.dumps(res))
