"""Synthetic AiMDoom-shaped scenes (host side, numpy).

The AiMDoom dataset is not in the reference tree (.gitignore:1-4, README.md:65) and there is no
network, so benchmarks and tests use seeded stand-ins with the same gross statistics
(SURVEY.md section 8d): unions of axis-aligned rooms joined on a lattice, walls / floor / ceiling
triangulated, already at the reference's x10 scene scale (configs/nbp/nbp_default_training_config.json:38),
room height 8-12 units, footprint 100-250 units, a triangle budget per difficulty level
(simple 4k / normal 12k / hard 25k / insane 50k).  Camera poses live on the reference's 3-unit
lattice at floor + 3.3 (macarons_utils.py:2301,2317-2319) with 8 azimuths (45 degree steps).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

TRI_BUDGET = {"simple": 4000, "normal": 12000, "hard": 25000, "insane": 50000}
CELL = 12.0          # room-cell edge in scene units: 4 camera-lattice steps of 3 units
LATTICE = 3.0
CAM_HEIGHT = 3.3


@dataclass
class Scene:
    verts: np.ndarray        # (V, 3) float32
    faces: np.ndarray        # (F, 3) int32
    open_cells: np.ndarray   # (G, G) bool, [ix, iz]
    floor_y: float
    ceil_y: float
    origin: np.ndarray       # (2,) world x,z of cell (0,0)'s low corner

    @property
    def n_faces(self) -> int:
        return int(self.faces.shape[0])


def _quad(p00, p10, p11, p01, n, verts, faces):
    """Append an n x n grid of sub-quads (2 triangles each) spanning the bilinear patch."""
    base = sum(len(v) for v in verts)
    u = np.linspace(0.0, 1.0, n + 1, dtype=np.float64)
    uu, vv = np.meshgrid(u, u, indexing="ij")
    P = ((1 - uu)[..., None] * (1 - vv)[..., None] * p00 + uu[..., None] * (1 - vv)[..., None] * p10
         + uu[..., None] * vv[..., None] * p11 + (1 - uu)[..., None] * vv[..., None] * p01)
    verts.append(P.reshape(-1, 3))
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1) + base
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    faces.append(np.stack([a, b, c], 1))
    faces.append(np.stack([a, c, d], 1))


def make_scene(seed: int, level: str = "simple", grid: int | None = None, tri_budget: int | None = None) -> Scene:
    """One seeded scene.  ``grid`` x ``grid`` cells of 12 units; a random connected subset is open."""
    rng = np.random.default_rng(seed)
    budget = int(tri_budget if tri_budget is not None else TRI_BUDGET[level])
    G = int(grid if grid is not None else rng.integers(9, 15))          # footprint 108-168 units
    n_open = int(G * G * rng.uniform(0.35, 0.55))
    open_cells = np.zeros((G, G), dtype=bool)
    cur = (G // 2, G // 2)
    open_cells[cur] = True
    frontier = [cur]
    while open_cells.sum() < n_open:                                     # randomised growth: rooms + corridors
        cx, cz = frontier[rng.integers(len(frontier))]
        dx, dz = [(1, 0), (-1, 0), (0, 1), (0, -1)][rng.integers(4)]
        run = int(rng.integers(1, 4))
        for _ in range(run):
            cx, cz = cx + dx, cz + dz
            if not (0 <= cx < G and 0 <= cz < G):
                break
            if not open_cells[cx, cz]:
                open_cells[cx, cz] = True
                frontier.append((cx, cz))
    floor_y = float(rng.uniform(-2.0, 2.0))
    ceil_y = floor_y + float(rng.uniform(8.0, 12.0))
    origin = np.array([-G * CELL / 2, -G * CELL / 2]) + rng.uniform(-5, 5, size=2)

    # count quads, then pick the subdivision that meets the triangle budget
    quads = []
    for ix in range(G):
        for iz in range(G):
            if not open_cells[ix, iz]:
                continue
            x0, z0 = origin[0] + ix * CELL, origin[1] + iz * CELL
            x1, z1 = x0 + CELL, z0 + CELL
            f, c = floor_y, ceil_y
            quads.append(([x0, f, z0], [x1, f, z0], [x1, f, z1], [x0, f, z1]))       # floor
            quads.append(([x0, c, z0], [x0, c, z1], [x1, c, z1], [x1, c, z0]))       # ceiling
            for (nx_, nz_, a, b) in ((ix - 1, iz, (x0, z0), (x0, z1)), (ix + 1, iz, (x1, z1), (x1, z0)),
                                     (ix, iz - 1, (x1, z0), (x0, z0)), (ix, iz + 1, (x0, z1), (x1, z1))):
                closed = not (0 <= nx_ < G and 0 <= nz_ < G) or not open_cells[nx_, nz_]
                if closed:
                    quads.append(([a[0], f, a[1]], [b[0], f, b[1]], [b[0], c, b[1]], [a[0], c, a[1]]))
    n = max(1, int(round(np.sqrt(budget / (2.0 * len(quads))))))
    verts, faces = [], []
    for q in quads:
        _quad(*[np.asarray(p, dtype=np.float64) for p in q], n, verts, faces)
    V = np.concatenate(verts).astype(np.float32)
    Fc = np.concatenate(faces).astype(np.int32)
    return Scene(V, Fc, open_cells, floor_y, ceil_y, origin.astype(np.float64))


def lattice_positions(scene: Scene) -> np.ndarray:
    """All camera lattice positions (x, y, z) inside open cells, 1.5 units clear of every wall."""
    G = scene.open_cells.shape[0]
    out = []
    offs = np.arange(LATTICE / 2, CELL, LATTICE)
    for ix in range(G):
        for iz in range(G):
            if scene.open_cells[ix, iz]:
                for ox in offs:
                    for oz in offs:
                        out.append((scene.origin[0] + ix * CELL + ox, scene.floor_y + CAM_HEIGHT,
                                    scene.origin[1] + iz * CELL + oz))
    return np.asarray(out, dtype=np.float32)


def _is_open(scene: Scene, x: float, z: float) -> bool:
    G = scene.open_cells.shape[0]
    ix = int(np.floor((x - scene.origin[0]) / CELL))
    iz = int(np.floor((z - scene.origin[1]) / CELL))
    return 0 <= ix < G and 0 <= iz < G and bool(scene.open_cells[ix, iz])


def random_walk(scene: Scene, n_poses: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    """A seeded key-pose trajectory on the lattice: each move is one 3-unit step along x or z, or a
    45 degree turn in place.  Returns poses (n_poses, 5) = (x, y, z, elev, azim deg) and the azimuth
    indices (n_poses,) in 0..7 (the reference's wrap logic at macarons_utils.py:2616-2621 needs them)."""
    rng = np.random.default_rng(seed)
    pos = lattice_positions(scene)
    p = pos[rng.integers(len(pos))].astype(np.float64)
    az = int(rng.integers(8))
    poses = np.zeros((n_poses, 5), dtype=np.float32)
    az_idx = np.zeros(n_poses, dtype=np.int64)
    for t in range(n_poses):
        poses[t] = (p[0], p[1], p[2], 0.0, 360.0 * az / 8)
        az_idx[t] = az
        for _ in range(16):
            kind = rng.integers(3)
            if kind == 0:
                az = (az + (1 if rng.integers(2) else -1)) % 8
                break
            d = LATTICE * (1 if rng.integers(2) else -1)
            q = p.copy()
            q[0 if kind == 1 else 2] += d
            if _is_open(scene, q[0], q[2]):
                p = q
                break
    return poses, az_idx


def sample_surface(scene: Scene, n_points: int, seed: int) -> np.ndarray:
    """Area-weighted random points on the mesh surface: the ground-truth cloud the coverage metric compares against
    (the reference samples `n_gt_surface_points` = 20000 / 50000 of them per scene, macarons_utils.py:612-637)."""
    rng = np.random.default_rng(seed)
    tri = scene.verts[scene.faces]                                   # (F, 3, 3)
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    f = rng.choice(len(tri), size=n_points, p=area / area.sum())
    u, v = rng.random(n_points), rng.random(n_points)
    flip = u + v > 1.0
    u[flip], v[flip] = 1.0 - u[flip], 1.0 - v[flip]
    p = tri[f, 0] + u[:, None] * (tri[f, 1] - tri[f, 0]) + v[:, None] * (tri[f, 2] - tri[f, 0])
    return p.astype(np.float32)


# ------------------------------------------------------------------------------------------------ seeded network weights / inputs
def count_like_input(B: int, S: int, seed: int = 8, density: float = 0.04, rate: float = 25.0, n_traj: int = 20):
    """Synthetic model input shaped like nbp_planning.py:126-132: 4 sparse count images + a sparse 0/1 trajectory image
    (integer-valued fp32, host tensor).  Same recipe (and, for a seed, the same values) as the oracle's generator."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(B, 5, S, S)
    occ = (torch.rand(B, 4, S, S, generator=g) < density).float()
    x[:, :4] = occ * (1.0 + torch.poisson(torch.full((B, 4, S, S), rate), generator=g))
    idx = torch.randint(0, S, (B, n_traj, 2), generator=g)
    for b in range(B):
        x[b, 4, idx[b, :, 0], idx[b, :, 1]] = 1.0
    return x


def seeded_nbp_state_dict(net, seed: int = 9):
    """Deterministic, well-conditioned weights for an ``NBP`` module, independent of torch's initialisers: conv weights
    U(-b, b) with b = sqrt(3 / fan_in), small biases, BatchNorm gamma in [0.8, 1.2], beta in [-0.1, 0.1], running_mean in
    [-0.2, 0.2], running_var in [0.6, 1.4]; drawn in state_dict order from one generator (there are no checkpoints offline:
    README.md:71-80 links them)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    u = lambda shape, lo, hi: torch.rand(tuple(shape), generator=g, dtype=torch.float32) * (hi - lo) + lo
    sd = {}
    for key, ref in net.state_dict().items():
        if key == "log_vars":
            sd[key] = torch.zeros_like(ref, device="cpu")
        elif key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(0, dtype=torch.long)
        elif key.endswith("running_mean"):
            sd[key] = u(ref.shape, -0.2, 0.2)
        elif key.endswith("running_var"):
            sd[key] = u(ref.shape, 0.6, 1.4)
        elif ref.dim() == 4:
            fan_in = ref.shape[1] * ref.shape[2] * ref.shape[3]
            b = (3.0 / fan_in) ** 0.5
            sd[key] = u(ref.shape, -b, b)
        elif key.endswith(".weight"):                       # BatchNorm gamma (conv weights are 4-D)
            sd[key] = u(ref.shape, 0.8, 1.2)
        else:                                               # ".bias": conv bias if the sibling weight is 4-D, else BatchNorm beta
            conv = net.state_dict()[key[:-4] + "weight"].dim() == 4
            sd[key] = u(ref.shape, -0.05, 0.05) if conv else u(ref.shape, -0.1, 0.1)
    return sd


def calibrated_nbp(device, seed: int = 9, calib_S: int = 64, calib_B: int = 2, calib_x=None):
    """An eval-mode ``NBP`` on ``device`` with seeded weights whose BatchNorm running statistics are the batch statistics of
    count-like inputs (one train-mode pass with momentum 1 on the CUDA kernels) and whose value head is scaled so that the
    value map is O(1-10), like the x100 coverage gains the reference trains on (nbp_utils.py:668).  With default running
    statistics activations explode / vanish and every accuracy figure would be meaningless.
    ``calib_x`` (B,5,S,S): calibrate on these model inputs instead (e.g. grids of the workload about to be run -- what training on
    that data would leave in the running statistics)."""
    import torch
    from .networks import NBP
    net = NBP()
    net.load_state_dict(seeded_nbp_state_dict(net, seed))
    net.to(device)
    net.train()
    net.bn_momentum = 1.0
    with torch.no_grad():
        net(count_like_input(calib_B, calib_S, seed=seed + 1).to(device) if calib_x is None else calib_x.to(device))
    net.bn_momentum = 0.1
    net.eval()
    with torch.no_grad():
        for k, v in net.state_dict().items():
            if k.endswith("num_batches_tracked"):
                v.zero_()
        o1, _ = net(count_like_input(1, calib_S, seed=seed + 2).to(device) if calib_x is None else calib_x[:1].to(device))
        scale = 5.0 / float(o1.abs().max().clamp_min(1e-6))
        net.Final1.weight.mul_(scale)
        net.Final1.bias.mul_(scale)
    return net
