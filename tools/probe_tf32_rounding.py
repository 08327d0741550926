import sys, torch
sys.path.insert(0, '/root/repo')
from poissonvae_b200 import _lib
lib = _lib.load()
s = torch.cuda.current_stream().cuda_stream
M = N = 128; Kd = 32
for name, v in (("1+3*2^-12", 1 + 3 * 2.0 ** -12), ("1+2^-11+2^-23", 1 + 2.0 ** -11 + 2.0 ** -23), ("1+2^-10-2^-23", 1 + 2.0 ** -10 - 2.0 ** -23), ("-(1+3*2^-12)", -(1 + 3 * 2.0 ** -12))):
    A = torch.full((M, Kd), v, device="cuda"); B = torch.ones(N, Kd, device="cuda"); C = torch.empty(M, N, device="cuda")
    assert lib.pvae_gemm_tf32(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, Kd, 1, 1, 0, 1, None, s) == 0
    torch.cuda.synchronize()
    print(name, "A as tf32 ->", C[0, 0].item() / Kd, " trunc would be", float(torch.tensor(v).view(torch.int32).bitwise_and(-8192).view(torch.float32)))
