"""Seeded synthetic scenes of the reference's input shape.

The reference feeds ScanNet scans sub-sampled to 50 000 points (src/visual_data_handlers.py:84-126,
seed 1184) as (B, N, 6) float32 = xyz + mean-centred rgb (src/joint_det_dataset.py:83,479).  No
dataset can be shipped, so the tests and bench.py use these generators.  All generation happens on
the CPU with an explicit torch.Generator so the same seed gives the same bytes everywhere.

Families (xyz):
  "uniform"  uniform in an 8 x 6 x 3 m room [-4,4]x[-3,3]x[0,3]; the origin lies inside the room, so
             the reference's |p|^2 <= 1e-3 skip in FPS (sampling_gpu.cu:105-106) is exercised
  "surface"  points on the room's 6 faces and on the faces of 32 random cuboids, randomly permuted
             (ScanNet-like density: ~60 neighbours within r = 0.2 at N = 50 000)
  "dup"      "surface" with 10 % of the points replaced by copies of other points (scans with fewer
             than N vertices are sampled with replacement, visual_data_handlers.py:114-118): exact
             distance ties
  "lattice"  points on a 0.25 m grid, heavily duplicated: every FPS decision is a tie
  "origin"   "uniform" with 5 % of the points pulled inside |p| < 0.04 (skipped by FPS)
"""
import torch

SEED = 1184
ROOM_LO = torch.tensor([-4.0, -3.0, 0.0])
ROOM_HI = torch.tensor([4.0, 3.0, 3.0])
FAMILIES = ("uniform", "surface", "dup", "lattice", "origin")


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def _uniform(n, g):
    return ROOM_LO + (ROOM_HI - ROOM_LO) * torch.rand(n, 3, generator=g)


def _on_box_faces(lo, hi, n, g):
    """n points uniform on the 6 faces of the axis-aligned box [lo, hi] (area-weighted)."""
    ext = hi - lo
    areas = torch.stack([ext[1] * ext[2], ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[2], ext[0] * ext[1],
                         ext[0] * ext[1]])
    face = torch.multinomial(areas / areas.sum(), n, replacement=True, generator=g)
    p = lo + ext * torch.rand(n, 3, generator=g)
    axis = face // 2
    side = (face % 2).to(torch.float32)
    fixed = lo[axis] + ext[axis] * side
    p[torch.arange(n), axis] = fixed
    return p


def _surface(n, g):
    n_room = n // 2
    parts = [_on_box_faces(ROOM_LO, ROOM_HI, n_room, g)]
    n_obj = 32
    per = (n - n_room) // n_obj
    for i in range(n_obj):
        size = 0.2 + 1.8 * torch.rand(3, generator=g)
        centre = ROOM_LO + (ROOM_HI - ROOM_LO) * torch.rand(3, generator=g)
        lo = torch.maximum(centre - size / 2, ROOM_LO)
        hi = torch.minimum(centre + size / 2, ROOM_HI)
        cnt = per if i < n_obj - 1 else n - n_room - per * (n_obj - 1)
        parts.append(_on_box_faces(lo, hi, cnt, g))
    p = torch.cat(parts, 0)
    return p[torch.randperm(n, generator=g)]


def scene_xyz(n, family="surface", seed=SEED):
    """(n, 3) float32 CPU tensor."""
    g = _gen(seed)
    if family == "uniform":
        p = _uniform(n, g)
    elif family == "surface":
        p = _surface(n, g)
    elif family == "dup":
        p = _surface(n, g)
        k = max(1, n // 10)
        dst = torch.randperm(n, generator=g)[:k]
        src = torch.randint(0, n, (k,), generator=g)
        p[dst] = p[src].clone()
    elif family == "lattice":
        cells = ((ROOM_HI - ROOM_LO) / 0.25).to(torch.int64) + 1
        ijk = torch.stack([torch.randint(0, int(c), (n,), generator=g) for c in cells], 1)
        p = ROOM_LO + 0.25 * ijk.to(torch.float32)
    elif family == "origin":
        p = _uniform(n, g)
        k = max(1, n // 20)
        dst = torch.randperm(n, generator=g)[:k]
        p[dst] = (torch.rand(k, 3, generator=g) - 0.5) * 0.04
    else:
        raise ValueError(f"unknown family {family!r}")
    return p.to(torch.float32).contiguous()


def point_clouds(batch, n, family="surface", seed=SEED, channels=3):
    """(batch, n, 3 + channels) float32 CPU tensor: xyz + colour-like features.

    rgb ~ U[0,1) minus the reference's mean colour (109.8, 97.2, 83.8)/256, src/joint_det_dataset.py:83,479."""
    g = _gen(seed + 7919)
    xyz = torch.stack([scene_xyz(n, family, seed + 1000 * b) for b in range(batch)], 0)
    if channels == 0:
        return xyz
    feats = torch.rand(batch, n, channels, generator=g)
    if channels == 3:
        feats = feats - torch.tensor([109.8, 97.2, 83.8]) / 256.0
    return torch.cat([xyz, feats], -1).contiguous()
