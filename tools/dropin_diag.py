import sys, os, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import test_gpu_dropin as T
from deepsir_b200 import patch as P
from oracle import deepsir_oracle as O
M, args = T._reference()
for n, seed in ((4096, 0), (2048, 0), (4096, 1)):
    net = T._net(M, args, seed)
    data = T._data(n)
    with torch.no_grad():
        P.unpatch()
        tr0, ep0 = net(dict(data), (2, False))
        f0, x0, l0, s0, f1, x1, l1, s1 = net.forward_pair(dict(data))
        fs, fr = net.aggregation(x0, x1, f0, f1, l0, l1, s0, s1)
        P.patch()
        tr1, ep1 = net(dict(data), (2, False))
        P.unpatch()
    i64, gap = O.match_top2_fp64(fs.cpu(), fr.cpu())
    p0, p1 = ep0["pred_pairs"][0][0, :, 1], ep1["pred_pairs"][0][0, :, 1]
    w0 = ep0["perm_matrices"][0].sigmoid()
    print(f"n={n} seed={seed}: gap median {gap.median().item():.3e} min {gap.min().item():.3e}, ambiguous(<2e-6) {(gap < 2e-6).float().mean().item():.4f}, "
          f"rows differing it0 {(p0 != p1).sum().item()}, stock==fp64 {(p0 == i64[0].int()).float().mean().item():.4f}, lib==fp64 {(p1 == i64[0].int()).float().mean().item():.4f}, "
          f"weights min/max {w0.min().item():.4f}/{w0.max().item():.4f}, logit maxdiff {(ep0['perm_matrices'][0]-ep1['perm_matrices'][0]).abs().max().item():.3e}, "
          f"feat norm {fs.norm(dim=1).mean().item():.3f} feat std over points {fs.std(dim=2).mean().item():.3e}")
    ang = O.rotation_angle_deg(tr1[0][:, :, :3].cpu(), tr0[0][:, :, :3].cpu()).max().item()
    print("   pose diff it0 deg", ang, "dt", (tr1[0][:, :, 3] - tr0[0][:, :, 3]).norm(dim=1).max().item())
