"""Seeded synthetic inputs shaped like the reference's datasets.

Generated on the CPU in fp32 with ``torch.Generator().manual_seed`` so that the CUDA path, the
oracle and the golden fixtures all see identical bytes.  Shapes follow the loaders:
KITTI crop r in [3,60] m, z in [-3,10] (dataloader/kitti_loader.py:323); jitter N(0,0.01) clipped at
0.05 (dataloader/transformation.py:96-107); rotation/translation ranges from arguments.py:35-39;
3DMatch voxel 0.03 m rooms (dataloader/threeDMatch_loader.py:49,55); Oxford r<=50, z in [-3,20]
(dataloader/oxford_loader.py:168-169,35).
"""
from __future__ import annotations

import math

import torch

BASE_SEED = 20201023


def seed_for(config: int, pair: int = 0) -> int:
    return BASE_SEED + 1000 * config + pair


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


def _rot_zyx(yaw, pitch, roll):
    cy, sy, cp, sp, cr, sr = (math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch),
                              math.cos(roll), math.sin(roll))
    Rz = torch.tensor([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=torch.float64)
    Ry = torch.tensor([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=torch.float64)
    Rx = torch.tensor([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=torch.float64)
    return Rz @ Ry @ Rx


def random_pose(g, yaw_deg=45.0, tilt_scale=0.1, trans_mag=2.0, any_axis=False):
    u = torch.rand(6, generator=g, dtype=torch.float64)
    if any_axis:
        ang = [(2 * u[i].item() - 1) * math.radians(yaw_deg) for i in range(3)]
    else:
        ang = [(2 * u[0].item() - 1) * math.radians(yaw_deg),
               (2 * u[1].item() - 1) * math.radians(yaw_deg) * tilt_scale,
               (2 * u[2].item() - 1) * math.radians(yaw_deg) * tilt_scale]
    R = _rot_zyx(*ang)
    t = (2 * u[3:6] - 1)
    t = t / t.norm().clamp_min(1e-9) * trans_mag * torch.rand(1, generator=g, dtype=torch.float64)
    return torch.cat([R, t[:, None]], dim=1).float()  # [3,4]


def kitti_cloud(n, g, r_min=3.0, r_max=60.0):
    """LiDAR-like 2.5-D cloud [n,4]: xyz + reflectance."""
    r = r_min + (r_max - r_min) * torch.rand(n, generator=g)
    th = 2 * math.pi * torch.rand(n, generator=g)
    ground = torch.rand(n, generator=g) < 0.7
    zg = -1.7 + 0.05 * torch.randn(n, generator=g)
    zs = -1.7 + 7.7 * torch.rand(n, generator=g)
    z = torch.where(ground, zg, zs).clamp(-3.0, 10.0)
    refl = 0.99 * torch.rand(n, generator=g)
    return torch.stack([r * torch.cos(th), r * torch.sin(th), z, refl], dim=1)


def room_cloud(n, g, size=(3.0, 3.0, 2.5)):
    """3DMatch-like fragment: points on the six faces of a box room + 5 mm noise, [n,3]."""
    face = torch.randint(0, 6, (n,), generator=g)
    uvw = torch.rand(n, 3, generator=g)
    ax = face // 2
    side = (face % 2).float()
    uvw[torch.arange(n), ax] = side
    pts = uvw * torch.tensor(size)
    return pts + 0.005 * torch.randn(n, 3, generator=g)


def oxford_cloud(n, g):
    r = 50.0 * torch.sqrt(torch.rand(n, generator=g))
    th = 2 * math.pi * torch.rand(n, generator=g)
    ground = torch.rand(n, generator=g) < 0.6
    z = torch.where(ground, -2.0 + 0.1 * torch.randn(n, generator=g), -3.0 + 23.0 * torch.rand(n, generator=g))
    return torch.stack([r * torch.cos(th), r * torch.sin(th), z.clamp(-3.0, 20.0)], dim=1)


def planted_features(n, d, perm, g, noise=0.1, outlier_frac=0.1):
    """Unit features with planted matches: f_ref[perm[j]] = normalize(g_j + noise * n_j); a fraction of
    source rows get independent features (outliers).  Returns (f_src [d,n], f_ref [d,n], inlier [n] bool),
    channel-major like the reference's [B,C,N] tensors (network/model.py:233-234 normalises them)."""
    base = torch.randn(n, d, generator=g)
    f_src = torch.nn.functional.normalize(base, dim=1)
    f_ref_src_order = torch.nn.functional.normalize(base + noise * torch.randn(n, d, generator=g), dim=1)
    inlier = torch.rand(n, generator=g) >= outlier_frac
    rnd = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    f_src = torch.where(inlier[:, None], f_src, rnd)
    f_ref = torch.empty_like(f_ref_src_order)
    f_ref[perm] = f_ref_src_order
    return f_src.t().contiguous(), f_ref.t().contiguous(), inlier


def make_pair(n, d=64, kind="kitti", seed=0, tiled_frac=0.0):
    """One registration pair.  Returns a dict of CPU fp32 tensors:
    points_src/points_ref [n,C>=3], feat_src/feat_ref [d,n], weights [n,1], perm [n] (src j <-> ref perm[j]),
    transform_gt [3,4], inlier [n]."""
    g = _gen(seed)
    if kind == "kitti":
        src = kitti_cloud(n, g)
        T = random_pose(g, 45.0, 0.1, 2.0)
    elif kind == "3dmatch":
        src = room_cloud(n, g)
        T = random_pose(g, 90.0, 1.0, 0.5, any_axis=True)
    elif kind == "oxford":
        src = oxford_cloud(n, g)
        T = random_pose(g, 30.0, 0.1, 2.0)
    else:
        raise ValueError(kind)
    if tiled_frac > 0:  # FixedResampler-style exact duplicates (dataloader/transformation.py:83-93)
        m = int(n * tiled_frac)
        src[n - m:] = src[:m]
    xyz = src[:, :3]
    ref_xyz = xyz @ T[:, :3].t() + T[:, 3]
    jit = (0.01 * torch.randn(n, 3, generator=g)).clamp(-0.05, 0.05)
    ref_xyz = ref_xyz + jit
    perm = torch.randperm(n, generator=g)
    ref = src.clone()
    ref[:, :3] = ref_xyz
    ref_perm = torch.empty_like(ref)
    ref_perm[perm] = ref
    f_src, f_ref, inlier = planted_features(n, d, perm, g)
    logit = torch.where(inlier, 2.0 + torch.randn(n, generator=g), -2.0 + torch.randn(n, generator=g))
    return dict(points_src=src.contiguous(), points_ref=ref_perm.contiguous(), feat_src=f_src, feat_ref=f_ref,
                weights=torch.sigmoid(logit)[:, None].contiguous(), perm=perm, transform_gt=T, inlier=inlier)


def make_batch(batch, n, d=64, kind="kitti", config=2, first_pair=0, tiled_frac=0.0):
    """Stack `batch` pairs: points [B,n,C], feats [B,d,n], weights [B,n,1], perm [B,n], transform_gt [B,3,4]."""
    pairs = [make_pair(n, d, kind, seed_for(config, first_pair + i), tiled_frac) for i in range(batch)]
    return {k: torch.stack([p[k] for p in pairs]) for k in pairs[0]}


def random_features(batch, d, n, seed):
    """Unit features WITHOUT planted matches (small top-2 gaps; stresses the filter-and-refine argmin)."""
    g = _gen(seed)
    return torch.nn.functional.normalize(torch.randn(batch, d, n, generator=g), dim=1)
