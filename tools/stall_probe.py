"""Is the box quiet?  Launch tiny kernels + small synchronous D2H reads for ~12 s (no code of this
repo involved) and report every iteration that took > 5 ms, with its timestamp."""
import time, torch
x = torch.zeros(1 << 20, device='cuda')
torch.cuda.synchronize()
t0 = time.perf_counter(); last = t0; slow = []; n = 0
while time.perf_counter() - t0 < 12.0:
  for _ in range(20):
    x.add_(1.0)
  v = x[:4].cpu()
  now = time.perf_counter(); n += 1
  if now - last > 0.005:
    slow.append((round(now - t0, 2), round((now - last) * 1e3, 1)))
  last = now
print('iterations', n, 'mean ms', round(12e3 / n, 3))
print('slow iterations (t [s], ms):', slow[:60])
