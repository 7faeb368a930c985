"""GPU bring-up probe: runs one named check per process (so a hung kernel can be killed by
`timeout` without taking the rest down) and prints PASS/FAIL lines.  Dev tool, not a test."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from transfusion_b200 import ops


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def gemm_case(M, N, K, a_mn, b_mn, tile_n=0, split_k=1, f32=False, bias=False, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    ref = A.float() @ B.float().t()
    bias_t = torch.randn(N, device=dev) if bias else None
    if bias:
        ref = ref + bias_t
    a_st = A.t().contiguous() if a_mn else A
    b_st = B.t().contiguous() if b_mn else B
    if f32 or split_k > 1:
        out = torch.zeros(M, N, device=dev, dtype=torch.float32)
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(a_st, b_st, out, M=M, N=N, K=K, a_mn_major=a_mn, b_mn_major=b_mn, tile_n=tile_n,
             split_k=split_k, accumulate=split_k > 1, bias=bias_t)
    torch.cuda.synchronize()
    e = rel(out.float(), ref)
    tag = f"gemm M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} tile_n={tile_n} split={split_k} f32={int(f32)}"
    print(("PASS " if e < 6e-3 else "FAIL ") + tag + f" rel={e:.3e}", flush=True)
    if e >= 6e-3:
        d = (out.float() - ref).abs()
        bad = (d > 0.05 * ref.abs().max()).nonzero()
        print("   first bad idx:", bad[:8].tolist(), " n_bad:", bad.shape[0], flush=True)
        # error map per 32x32 block
        Mb, Nb = min(M, 256), min(N, 256)
        blk = d[:Mb, :Nb].reshape(Mb // 32, 32, Nb // 32, 32).amax(dim=(1, 3)) if Mb % 32 == 0 and Nb % 32 == 0 else None
        if blk is not None:
            print("   blockmax(32x32):\n", (blk > 0.05 * ref.abs().max()).int().cpu().numpy(), flush=True)


CASES = {
    "gemm_kk_small": lambda: gemm_case(128, 128, 64, False, False, tile_n=128),
    "gemm_kk_k256": lambda: gemm_case(128, 128, 256, False, False, tile_n=128),
    "gemm_kk_multi": lambda: gemm_case(1024, 896, 896, False, False, bias=True),
    "gemm_kk_tail": lambda: gemm_case(1000, 712, 712, False, False, bias=True),
    "gemm_kk_big": lambda: gemm_case(8192, 2688, 896, False, False),
    "gemm_kmn_small": lambda: gemm_case(128, 128, 64, False, True, tile_n=128),
    "gemm_kmn": lambda: gemm_case(512, 896, 1792, False, True),
    "gemm_mnk_small": lambda: gemm_case(128, 128, 64, True, False, tile_n=128),
    "gemm_mnmn_small": lambda: gemm_case(128, 128, 128, True, True, tile_n=128),
    "gemm_mnmn": lambda: gemm_case(896, 1792, 4096, True, True, split_k=4),
    "gemm_mnmn_tail": lambda: gemm_case(712, 1424, 1000, True, True, split_k=3),
    "gemm_f32": lambda: gemm_case(256, 256, 512, False, False, f32=True),
}

if __name__ == "__main__":
    name = sys.argv[1]
    if name == "list":
        print(" ".join(CASES))
    else:
        CASES[name]()
