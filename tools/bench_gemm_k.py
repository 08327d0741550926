"""GEMM time as a function of the reduction length: fixed cost vs per-K-block cost (M=4096, N=512)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poissonvae_b200 import _lib
from tools.bench_gemm import time_us

lib = _lib.load()
s = torch.cuda.current_stream().cuda_stream
M, N = 4096, 512
for precise in (0, 1):
    for Kd in (32, 64, 128, 256, 512, 1024):
        A = torch.randn(M, Kd, device="cuda"); B = torch.randn(N, Kd, device="cuda"); C = torch.empty(M, N, device="cuda")
        fn = lambda: lib.pvae_gemm_tf32(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, Kd, 1, 1, precise, 1, None, s)
        assert fn() == 0
        print(json.dumps(dict(precise=precise, Kd=Kd, us=round(time_us(fn), 2))), flush=True)
# an empty kernel launch for the event/launch overhead
t = time_us(lambda: torch.empty(1, device="cuda").zero_())
print(json.dumps(dict(torch_tiny_kernel_us=round(t, 2))))
