"""Top-k soft correspondences: fused two-sweep route against the materialising route (selected by a zero column bias) and
against the plain soft pass (no top-k), on BASELINE config-3-shaped inputs and a C2-width one.  Prints ms per call."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = "cuda:0"
    for (B, C, J, K, topk) in [(32, 32, 5000, 5000, 8), (32, 32, 5000, 5000, 32), (8, 64, 16384, 16384, 5), (8, 64, 16384, 16384, 32)]:
        g = torch.Generator().manual_seed(1)
        fs = torch.nn.functional.normalize(torch.randn(B, C, J, generator=g), dim=1).to(dev)
        fr = torch.nn.functional.normalize(torch.randn(B, C, K, generator=g), dim=1).to(dev)
        xyz = torch.randn(B, K, 3, generator=g).to(dev)
        beta, alpha = torch.full((B,), 10.0, device=dev), torch.full((B,), 0.5, device=dev)
        zero = torch.zeros(B, K, device=dev)
        t_soft = timed(lambda: D.match_soft(fs, fr, xyz, beta, alpha))
        t_fused = timed(lambda: D.match_soft(fs, fr, xyz, beta, alpha, topk=topk))
        t_mat = timed(lambda: D.match_soft(fs, fr, xyz, beta, alpha, col_bias=zero, topk=topk), reps=3)
        lib = D.lib()
        # rows that took the exhaustive pass (diagnostic; one more call through ctypes with a workspace we keep)
        import ctypes
        from deepsir_b200 import _lib as L
        (f1, _a), (f2, _b) = L.feat_cn(fs), L.feat_cn(fr)
        ws = L.workspace(lib.dsir_match_soft_topk_workspace_bytes(B, C, J, K, topk), fs.device)
        ti = torch.empty(B, J, topk, dtype=torch.int64, device=dev)
        tw = torch.empty(B, J, topk, device=dev)
        lse = torch.empty(B, J, device=dev)
        L.check(lib.dsir_match_soft(f1, f2, B, C, J, K, beta.data_ptr(), alpha.data_ptr(), None, None, None, lse.data_ptr(), topk,
                                    ti.data_ptr(), tw.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(fs.device)), "soft")
        ex = ctypes.c_int32(-1)
        lib.dsir_match_soft_topk_exhaustive_rows(ws.data_ptr(), ws.numel(), B, C, J, K, topk, ctypes.byref(ex), L.stream_ptr(fs.device))
        print(f"   exhaustive rows: {ex.value} of {B * J}")
        print(f"B{B} C{C} {J}x{K} k{topk}: soft only {t_soft:.3f} ms | + top-k fused {t_fused:.3f} ms | + top-k materialised {t_mat:.3f} ms"
              f" | workspace fused {lib.dsir_match_soft_topk_workspace_bytes(B, C, J, K, topk) / 2**20:.0f} MiB"
              f" (soft alone {lib.dsir_match_soft_workspace_bytes(B, C, J, K) / 2**20:.0f} MiB)")


if __name__ == "__main__":
    main()
