import time, torch, numpy as np
torch.manual_seed(0)
B = torch.randn(1, 126, 53, 53, dtype=torch.float64, device='cuda')
def tm(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
print('torch cuda eigvals 126x53x53: %.1f ms' % tm(lambda: torch.linalg.eigvals(B)))
Bc = B.cpu()
print('torch cpu eigvals: %.1f ms' % tm(lambda: torch.linalg.eigvals(Bc)))
print('d2h + cpu eigvals: %.1f ms' % tm(lambda: torch.linalg.eigvals(B.cpu())))
Bn = Bc.numpy()
print('numpy eigvals: %.1f ms' % tm(lambda: np.linalg.eigvals(Bn)))
B8 = torch.randn(8, 126, 53, 53, dtype=torch.float64, device='cuda')
print('torch cuda eigvals 8x126: %.1f ms' % tm(lambda: torch.linalg.eigvals(B8), 1))
print('d2h + cpu eigvals 8x126: %.1f ms' % tm(lambda: torch.linalg.eigvals(B8.cpu()), 1))
