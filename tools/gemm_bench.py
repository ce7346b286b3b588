"""Micro-benchmark of the gather-GEMM through the test hook: python tools/gemm_bench.py  (GPU box)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "voiceprintrecognition-paddlepaddle_b200"))
import torch
from ppvector import _lib

lib = _lib.load()
dev = torch.device("cuda:0")

def run(M, N, K, bn, bk, prec, iters=10):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5
    out = torch.empty(M, N, device=dev)
    nbytes = lib.ppv_gemm_test_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    def call():
        _lib.check(lib.ppv_gemm_test(_lib.ptr(A), _lib.ptr(W), None, None, None, 0, M, N, K, bn, bk, prec, _lib.ptr(out),
                                     C.c_void_p(ws.data_ptr()), nbytes, _lib.current_stream()), "gemm")
    call(); torch.cuda.synchronize()
    # the hook also converts A/W to planes each call: time that part separately with K tiny? -> subtract a conversion-only estimate
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000

if __name__ == "__main__":
    M = 78336
    for (N, K) in [(512, 512), (1536, 1536)]:
        for prec in (0, 1):
            for bn, bk in [(256, 64), (256, 32), (128, 64), (128, 32)]:
                us = run(M, N, K, bn, bk, prec)
                fl = 2.0 * M * N * K * (3 if prec == 0 else 1)
                print(f"N={N} K={K} prec={'x3' if prec==0 else 'x1'} BN={bn} BK={bk} nostore={os.environ.get('PPV_GEMM_NOSTORE','0')}: {us:8.1f} us (incl. operand conversion)  {fl/us/1e6:7.1f} TF/s executed", flush=True)
