import sys, torch
sys.path.insert(0, '/root/repo')
from yolat_vectorgraphicsrecognition_b200 import ops
dev = torch.device('cuda')
M, N, K = 20000, 1024, 128
a, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
out = torch.empty(M, N, device=dev)
for _ in range(3):
    ops.gemm(0, a, b, out=out)
a2, b2 = torch.randn(M, 1024, device=dev), torch.randn(1024, 128, device=dev)
out2 = torch.empty(M, 128, device=dev)
for _ in range(3):
    ops.gemm(1, a2, b2, out=out2)
torch.cuda.synchronize()
