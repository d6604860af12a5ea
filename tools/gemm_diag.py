"""GPU bring-up diagnostic for the tcgen05 GEMM: compares against torch.matmul and prints an error map."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticattack_b200 import _lib

def run(M, N, K, bias=False, act=0, gamma=False, resid=False, f32=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = (torch.randn(N, device="cuda", generator=g) * 0.1).bfloat16() if bias else None
    gm = (torch.randn(N, device="cuda", generator=g)).bfloat16() if gamma else None
    r = (torch.randn(M, N, device="cuda", generator=g)).bfloat16() if resid else None
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    L = _lib.lib()
    rc = L.vla_gemm_bf16_tn(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(out), N, M, N, K, _lib.ptr(b), _lib.ptr(gm),
                            _lib.ptr(r), N, act, None, int(f32), _lib.cur_stream())
    _lib.check(rc, "gemm")
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    if bias: ref = ref + b.float()
    ref = ref.bfloat16().float()
    if act == 1: ref = torch.nn.functional.gelu(ref).bfloat16().float()
    if gamma: ref = (ref * gm.float()).bfloat16().float()
    if resid: ref = (r.float() + ref).bfloat16().float()
    err = (out.float() - ref).abs()
    nan = torch.isnan(out.float()).sum().item()
    tol = 2e-2 * ref.abs().max().item() + 1e-3
    bad = (err > tol) | torch.isnan(err)
    print(f"M={M} N={N} K={K} bias={bias} act={act} gamma={gamma} resid={resid} f32={f32}: max_err={err[~torch.isnan(err)].max().item() if (~torch.isnan(err)).any() else float('nan'):.4g} "
          f"ref_max={ref.abs().max().item():.3g} nan={nan} bad={bad.sum().item()}/{bad.numel()}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten().tolist()
        cols = bad.any(0).nonzero().flatten().tolist()
        print("   bad rows (first 20):", rows[:20], "... count", len(rows))
        print("   bad cols (first 20):", cols[:20], "... count", len(cols))
        print("   out[0,:8]", out[0, :8].float().tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
    return not bad.any().item()

def bench(M, N, K, iters=20):
    A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L = _lib.lib()
    def call():
        _lib.check(L.vla_gemm_bf16_tn(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(out), N, M, N, K, None, None, None, 0, 0, None, 0, _lib.cur_stream()))
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    e0.record()
    for _ in range(iters): torch.matmul(A, W.t(), out=out)
    e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"bench M={M} N={N} K={K}: {ms*1e3:.1f} us  {tf:.1f} TFLOP/s   (cuBLAS {ms_ref*1e3:.1f} us {2.0*M*N*K/ms_ref/1e9:.1f} TFLOP/s)")

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    L = _lib.lib()
    allok = True
    for mode in [(1, 256), (1, 128), (2, 256), (2, 128), (0, 0)]:
        _lib.check(L.vla_gemm_set_mode(*mode))
        print("=== variant (ctas, block_n) =", mode)
        ok = True
        ok &= run(128, 128, 64)
        ok &= run(256, 256, 256)
        ok &= run(300, 432, 144)
        ok &= run(100, 200, 136)            # ragged everything
        ok &= run(2304, 4096, 4096, bias=True)
        ok &= run(2088, 4096, 1024, bias=True, act=1)
        ok &= run(2088, 1024, 4096, bias=True, gamma=True, resid=True)
        ok &= run(64, 32064, 4096, f32=True)
        ok &= run(2048, 1152, 640, bias=True)
        print("variant OK" if ok else "variant FAILED")
        allok &= ok
        if ok:
            for shp in [(2304, 4096, 4096), (2304, 12288, 4096), (2304, 22016, 4096), (2304, 4096, 11008), (2088, 3072, 1024),
                        (2088, 1024, 4096), (2048, 4304, 1152), (8192, 8192, 8192)]:
                bench(*shp)
    print("ALL OK" if allok else "FAILURES")
