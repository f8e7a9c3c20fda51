"""Time the tcgen05 GEMM on the ViT layer shapes (development aid): full kernel, without global stores, mainloop only."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upliftingtabletennis_b200 import _lib
from upliftingtabletennis_b200._lib import lib, check, ptr, stream_ptr
dev = torch.device('cuda')
M = 23040
for name, N, K, act, res, c_bf16 in (('qkv', 1152, 384, 0, False, True), ('proj', 384, 384, 0, True, False), ('fc1', 1536, 384, 1, False, True),
                                     ('fc2', 384, 1536, 0, True, False)):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    R = torch.randn(M, N, device=dev) if res else None
    C = torch.empty((M, N), dtype=torch.bfloat16 if c_bf16 else torch.float32, device=dev)
    for dbg, label in ((0, 'full'), (2, 'no global stores'), (1, 'mainloop only')):
        a = act | (dbg << 8)
        for _ in range(3):
            check(lib.ttk_vit_debug_gemm(ptr(A), ptr(W), ptr(bias), ptr(R) if res else None, ptr(C), M, N, K, a, int(c_bf16), 0, 0, 0, 0, _lib.BF16, stream_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            check(lib.ttk_vit_debug_gemm(ptr(A), ptr(W), ptr(bias), ptr(R) if res else None, ptr(C), M, N, K, a, int(c_bf16), 0, 0, 0, 0, _lib.BF16, stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print('%-5s M %d N %4d K %4d  %-18s %7.1f us  %6.1f TFLOP/s' % (name, M, N, K, label, us, 2.0 * M * N * K / us / 1e6))
