"""Probe: does the tensor core / TMA round or truncate fp32 operands fed to tcgen05 kind::tf32?
Run twice: NRMS_TMA_TF32_DTYPE=0 (FLOAT32 tensor map) and =1 (TFLOAT32 tensor map)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from newsrecommendationsystem_b200 import ops, _lib

def rna_tf32(x):
    b = x.view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)

def trunc_tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)

torch.manual_seed(0)
dev = "cuda"
M, N, K = 4096, 900, 300
a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev) * 0.1
ref = a.double() @ b.double().T
def err(c, r=ref): return ((c.double() - r).abs().max() / r.abs().max()).item()
c_raw = ops.gemm_nt(a, b, None, mode=_lib.MODE_TF32)
c_rna = ops.gemm_nt(rna_tf32(a), rna_tf32(b), None, mode=_lib.MODE_TF32)
print("dtype flag", os.environ.get("NRMS_TMA_TF32_DTYPE", "0"))
print("raw inputs      vs fp64:", err(c_raw))
print("rna inputs      vs fp64:", err(c_rna))
print("raw vs exact-trunc model:", err(c_raw, trunc_tf32(a).double() @ trunc_tf32(b).double().T))
print("raw vs exact-rna   model:", err(c_raw, rna_tf32(a).double() @ rna_tf32(b).double().T))
# timing
for (m, n, k) in [(40960, 900, 300), (40960, 200, 300), (140800, 900, 300)]:
    a = torch.randn(m, k, device=dev); b = torch.randn(n, k, device=dev)
    for _ in range(3): ops.gemm_nt(a, b, None, mode=_lib.MODE_TF32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm_nt(a, b, None, mode=_lib.MODE_TF32)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"gemm {m}x{n}x{k}: {ms*1e3:.1f} us  {2*m*n*k/ms/1e9:.1f} TFLOP/s")
