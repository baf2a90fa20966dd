import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")   # 4 GB
y = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print("memset 4GB: %.3f ms  %.2f TB/s write" % (ms, 4.295 / ms))
ms = t(lambda: x.fill_(1.5)); print("fill   4GB: %.3f ms  %.2f TB/s write" % (ms, 4.295 / ms))
ms = t(lambda: y.copy_(x)); print("copy   4GB: %.3f ms  %.2f TB/s r+w" % (ms, 8.59 / ms))
ms = t(lambda: x.sum()); print("sum    4GB: %.3f ms  %.2f TB/s read" % (ms, 4.295 / ms))
h = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
ms = t(lambda: h.copy_(x)); print("cast f32->bf16: %.3f ms  %.2f TB/s r+w" % (ms, 6.44 / ms))
