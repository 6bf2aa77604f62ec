import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from deepsir_b200 import synth
from scipy.spatial import cKDTree
b = synth.make_batch(1, 16384, 8, "kitti", config=2)
P = b["points_src"][0, :, :3].numpy().astype(np.float64)
n = len(P); k = 16
tree = cKDTree(P)
dk, _ = tree.query(P, k=k)
rk = dk[:, -1]               # exact k-th distance per query

def morton_order(P):
    lo = P.min(0); ext = (P.max(0) - lo).max()
    q = np.clip(((P - lo) * (1023.0 / ext)).astype(np.int64), 0, 1023)
    def part(v):
        v = v & 0x3ff
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    key = part(q[:, 0]) | (part(q[:, 1]) << 1) | (part(q[:, 2]) << 2)
    return np.argsort(key, kind="stable")

def kd_order(P, leaf=32):
    idx = np.arange(len(P))
    out = []
    def rec(ix):
        if len(ix) <= leaf:
            out.append(ix); return
        pts = P[ix]
        ax = np.argmax(pts.max(0) - pts.min(0))
        # split at a multiple of leaf closest to the median
        m = (len(ix) // 2 + leaf - 1) // leaf * leaf
        o = np.argsort(pts[:, ax], kind="stable")
        rec(ix[o[:m]]); rec(ix[o[m:]])
    rec(idx)
    return np.concatenate(out)

def hilbert_like_xy_order(P):  # STR: x slabs then y
    nl = len(P) // 32
    s = int(round(np.sqrt(nl)))
    o = np.argsort(P[:, 0], kind="stable")
    out = []
    per = int(np.ceil(len(P) / s / 32)) * 32
    for i in range(0, len(P), per):
        sl = o[i:i + per]
        out.append(sl[np.argsort(P[sl, 1], kind="stable")])
    return np.concatenate(out)

def stats(order, name, leaf=32, gran=32):
    Q = P[order]; r = rk[order]
    nl = n // leaf
    ng = n // gran
    glo = Q.reshape(ng, gran, 3).min(1); ghi = Q.reshape(ng, gran, 3).max(1)
    tot_union = 0; tot_need = 0; tot_bbox = 0
    for L in range(nl):
        q = Q[L * leaf:(L + 1) * leaf]; rr = r[L * leaf:(L + 1) * leaf]
        # per query distance to each granule box
        g = np.maximum(np.maximum(glo[None] - q[:, None], q[:, None] - ghi[None]), 0)
        d = np.sqrt((g ** 2).sum(2))            # [leaf, ng]
        need = d <= rr[:, None]
        tot_union += need.any(0).sum()
        tot_need += need.sum() / leaf
        # coarse: qbox-to-box with max r
        qlo = q.min(0); qhi = q.max(0)
        gb = np.maximum(np.maximum(glo - qhi, qlo - ghi), 0)
        tot_bbox += (np.sqrt((gb ** 2).sum(1)) <= rr.max()).sum()
    print(f"{name:10s} group={leaf} granule={gran}: granules needed by any lane {tot_union / nl:6.1f} (= {tot_union / nl * gran:6.0f} cands/lane), per-lane need {tot_need / nl:5.1f} (= {tot_need / nl * gran:5.0f} cands), coarse-test passes {tot_bbox / nl:6.1f}")

for name, fn in (("morton", morton_order), ("kd", kd_order), ("str-xy", hilbert_like_xy_order)):
    o = fn(P)
    for gran in (32, 16, 8):
        stats(o, name, 32, gran)
    stats(o, name, 16, 16)

print("---- generic partition functions on other cloud shapes")
def str_adaptive(P, levels, leaf=32):
    """levels: list of fanouts; each level sorts every segment along its own widest axis and cuts it into `f` parts
    (boundaries at multiples of `leaf`)."""
    segs = [np.arange(len(P))]
    for f in levels:
        new = []
        for ix in segs:
            pts = P[ix]
            ax = np.argmax(pts.max(0) - pts.min(0))
            o = ix[np.argsort(pts[:, ax], kind="stable")]
            nl = (len(o) + leaf - 1) // leaf
            cuts = [min(len(o), ((j * nl) // f) * leaf) for j in range(f + 1)]
            cuts[-1] = len(o)
            for j in range(f):
                if cuts[j + 1] > cuts[j]:
                    new.append(o[cuts[j]:cuts[j + 1]])
        segs = new
    return np.concatenate(segs)

def run(Pnew, label):
    global P, n, rk
    P = Pnew.astype(np.float64); n = len(P)
    dk, _ = cKDTree(P).query(P, k=k)
    rk = dk[:, -1]
    print("==", label, n)
    stats(morton_order(P), "morton", 32, 32)
    stats(kd_order(P), "kd", 32, 32)
    nl = n // 32
    s2 = int(round(np.sqrt(nl)))
    stats(str_adaptive(P, [s2, (nl + s2 - 1) // s2]), f"str2 {s2}x{(nl + s2 - 1) // s2}", 32, 32)
    c = int(round(nl ** (1 / 3)))
    stats(str_adaptive(P, [c, c, (nl + c * c - 1) // (c * c)]), f"str3 {c}", 32, 32)
    stats(str_adaptive(P, [8, 8, (nl + 63) // 64]), "str3 8,8,x", 32, 32)

g = torch.Generator().manual_seed(1)
run(b["points_src"][0, :, :3].numpy(), "kitti 16384")
run(b["points_src"][0, :4096, :3].numpy(), "kitti 4096 (level 1)")
c3 = synth.make_batch(1, 5000, 8, "3dmatch", config=3)
run(c3["points_src"][0, :4992, :3].numpy(), "3dmatch 4992")
run(torch.rand(16384, 3, generator=g).numpy(), "uniform cube 16384")
run(torch.randn(8192, 3, generator=g).numpy(), "gaussian 8192")
