"""Seeded synthetic BAL-shaped problems in the reference's 15-parameter format.

The BAL files are not available offline, so every config of BASELINE.json is a
synthetic scene of the named shape:

  * cameras on a smooth forward-moving trajectory looking at a slab of points,
    pixel observations in the BAL convention (p = -P/P_z * f, y up);
  * landmark j is seen by deg_j cameras drawn from a window of nearby cameras
    (banded co-visibility, like a real sequence), with a small share of
    long-track landmarks so that landmark degrees have a tail;
  * Gaussian pixel noise.

The text file written here is what the reference's `--create-dataset`
(`/root/reference/src/rootba_povar/bal/bal_problem.cpp:306-471`) would leave in
`data_custom/`: header `C L N`, N lines `cam lm x y` (`%lf`, 6 decimals), then
15 numbers per camera -- the first two rows of the 3x4 camera matrix drawn from
N(0,1), the third row `0 0 0 1` (bal_problem.cpp:392-407), then f k1 k2 -- and 3
numbers per landmark (ignored by the loader, which redraws them;
bal_problem.cpp:255-268).  The reference seeds that draw from
std::random_device, so it is not reproducible; this generator draws the same
distribution from a numpy seed instead and both solvers read the same file.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

# name -> (cameras, landmarks, target observations); shapes from BASELINE.json
SHAPES = {
    "tiny": (6, 40, 160),
    "small": (16, 400, 1700),
    "ladybug49": (49, 7776, 31843),
    "trafalgar257": (257, 65132, 225911),
    "venice89": (89, 110973, 562976),
    "venice1778": (1778, 993923, 5001946),
    "scaled10k": (10000, 10000000, 50000000),
}


@dataclasses.dataclass
class SyntheticProblem:
    num_cams: int
    num_lms: int
    obs_cam: np.ndarray   # int32 [nnz]  file order: landmark-major, camera ascending
    obs_lm: np.ndarray    # int32 [nnz]
    obs_xy: np.ndarray    # float64 [nnz, 2]  BAL convention (y up), rounded to 6 decimals
    cam_params: np.ndarray  # float64 [C, 15]  rounded to 6 decimals
    points: np.ndarray    # float64 [L, 3]

    @property
    def num_obs(self) -> int:
        return int(self.obs_cam.shape[0])


def _sorted_unique(a: np.ndarray) -> np.ndarray:
    # np.unique(a) without numpy 2.3's hash-table path (10x slower than a sort at 5e7 keys)
    k = np.sort(a)
    if k.size == 0:
        return k
    keep = np.empty(k.shape[0], dtype=bool)
    keep[0] = True
    np.not_equal(k[1:], k[:-1], out=keep[1:])
    return k[keep]


def _round6(a: np.ndarray) -> np.ndarray:
    # value the reference reads back after `fprintf("%lf")`
    return np.round(a, 6)


def generate(num_cams: int, num_lms: int, target_obs: int, seed: int,
             noise_px: float = 0.5, focal: float = 1000.0,
             long_track_share: float = 0.01, arc: float = 2.0,
             extent: float = 2.0, window_frac: float = 0.25) -> SyntheticProblem:
    rng = np.random.default_rng(seed)
    C, L = int(num_cams), int(num_lms)
    mean_deg = max(2.0, target_obs / L)

    # ---- cameras: an arc around the scene, every camera looking at the centre
    # (BAL convention: the camera looks down its -z axis)
    k = np.arange(C, dtype=np.float64)
    theta = (k / max(C - 1, 1) - 0.5) * arc
    radius = 10.0
    centers = radius * np.stack([np.sin(theta), 0.15 * np.sin(3.0 * theta), np.cos(theta)], axis=1)
    centers = centers + rng.normal(0.0, 0.05, centers.shape)
    zc = centers / np.linalg.norm(centers, axis=1, keepdims=True)
    xc = np.cross(np.array([0.0, 1.0, 0.0])[None], zc)
    xc /= np.linalg.norm(xc, axis=1, keepdims=True)
    yc = np.cross(zc, xc)
    R = np.stack([xc, yc, zc], axis=1)               # rows = camera axes in world coordinates

    # ---- landmarks: a box at the centre of the arc
    pts = rng.uniform(-1.0, 1.0, (L, 3)) * np.array([extent, 0.6 * extent, extent])[None]

    # ---- visibility: degrees with a tail, cameras from a window around the
    # camera nearest to the landmark
    mult_mean = 7.5                                 # E[integers(4, 12)]
    mu = max(mean_deg / (1.0 + long_track_share * (mult_mean - 1.0)) - 2.0, 0.0)
    deg = 2 + rng.poisson(mu, L)
    long_tracks = rng.random(L) < long_track_share
    deg = np.where(long_tracks, deg * rng.integers(4, 12, L), deg)
    deg = np.clip(deg, 2, C).astype(np.int64)
    # window half-width in cameras; its centre is a random camera
    half = np.maximum(np.maximum(deg, 6), int(round(window_frac * C)))
    near = rng.integers(0, C, L)

    lm_rep = np.repeat(np.arange(L, dtype=np.int64), deg)
    # stratified offsets: the i-th pick of a landmark lands in the i-th slot of
    # its window, so picks are distinct before clipping
    pos_in_lm = np.arange(lm_rep.shape[0], dtype=np.int64) - np.repeat(
        np.cumsum(deg) - deg, deg)
    width = 2 * half[lm_rep] + 1
    slot = (pos_in_lm * width) // deg[lm_rep]
    slot_w = np.maximum(((pos_in_lm + 1) * width) // deg[lm_rep] - slot, 1)
    off = slot + (rng.random(lm_rep.shape[0]) * slot_w).astype(np.int64) - half[lm_rep]
    cam = near[lm_rep] + off
    # reflect at the ends of the trajectory instead of clipping (keeps picks distinct)
    cam = np.where(cam < 0, -cam - 1 + 0, cam)
    cam = np.where(cam > C - 1, 2 * (C - 1) - cam + 1, cam)
    cam = np.clip(cam, 0, C - 1)
    key = _sorted_unique(lm_rep * C + cam)               # sorted: landmark-major, camera ascending
    obs_lm = (key // C).astype(np.int64)
    obs_cam = (key % C).astype(np.int64)

    # every landmark needs >= 2 observations
    cnt = np.bincount(obs_lm, minlength=L)
    short = np.nonzero(cnt < 2)[0]
    if short.size:
        extra_lm, extra_cam = [], []
        have = {int(l): set() for l in short}
        sel = np.isin(obs_lm, short)
        for l, c in zip(obs_lm[sel], obs_cam[sel]):
            have[int(l)].add(int(c))
        for l in short:
            c0 = int(near[l])
            cand = [c0, min(c0 + 1, C - 1), max(c0 - 1, 0), min(c0 + 2, C - 1), max(c0 - 2, 0)]
            for c in cand:
                if len(have[int(l)]) >= 2:
                    break
                if c not in have[int(l)]:
                    have[int(l)].add(c)
                    extra_lm.append(int(l))
                    extra_cam.append(c)
        key = _sorted_unique(np.concatenate([key, np.asarray(extra_lm, np.int64) * C +
                                             np.asarray(extra_cam, np.int64)]))
        obs_lm = (key // C).astype(np.int64)
        obs_cam = (key % C).astype(np.int64)

    # ---- project (BAL / Snavely convention, no distortion)
    Pc = np.einsum("nij,nj->ni", R[obs_cam], pts[obs_lm] - centers[obs_cam])
    xy = -Pc[:, :2] / Pc[:, 2:3] * focal
    xy = xy + rng.normal(0.0, noise_px, xy.shape)

    # ---- randomised camera matrices, like --create-dataset
    cam_params = np.zeros((C, 15))
    cam_params[:, 0:8] = rng.normal(0.0, 1.0, (C, 8))
    cam_params[:, 11] = 1.0
    cam_params[:, 12] = focal

    return SyntheticProblem(C, L, obs_cam.astype(np.int32), obs_lm.astype(np.int32),
                            _round6(xy), _round6(cam_params), _round6(pts))


def generate_named(name: str, seed: int | None = None) -> SyntheticProblem:
    C, L, N = SHAPES[name]
    if seed is None:
        seed = 1000 + list(SHAPES).index(name)
    return generate(C, L, N, seed)


def write_bal(problem: SyntheticProblem, path: str, shuffle_seed: int | None = None) -> None:
    """Write the 15-parameter `data_custom` text file.

    `shuffle_seed` permutes the observation lines (the loader's canonical order
    is independent of file order) -- used by the indexing tests.
    """
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    order = np.arange(problem.num_obs)
    if shuffle_seed is not None:
        order = np.random.default_rng(shuffle_seed).permutation(problem.num_obs)
    with open(path, "w") as f:
        f.write(f"{problem.num_cams} {problem.num_lms} {problem.num_obs}")
        cam = problem.obs_cam[order]
        lm = problem.obs_lm[order]
        xy = problem.obs_xy[order]
        chunk = 1 << 20
        for s in range(0, problem.num_obs, chunk):
            e = min(s + chunk, problem.num_obs)
            lines = np.char.add(
                np.char.add(np.char.add(cam[s:e].astype(str), " "),
                            np.char.add(lm[s:e].astype(str), " ")),
                np.char.add(np.char.add(np.char.mod("%.6f", xy[s:e, 0]), " "),
                            np.char.mod("%.6f", xy[s:e, 1])))
            f.write("\n" + "\n".join(lines.tolist()))
        f.write("\n" + "\n".join("%.6f" % v for v in problem.cam_params.reshape(-1)))
        f.write("\n" + "\n".join("%.6f" % v for v in problem.points.reshape(-1)))
        f.write("\n")


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("name", choices=list(SHAPES))
    ap.add_argument("out")
    ap.add_argument("--seed", type=int, default=None)
    a = ap.parse_args()
    p = generate_named(a.name, a.seed)
    write_bal(p, a.out)
    print(f"{a.name}: {p.num_cams} cams, {p.num_lms} lms, {p.num_obs} obs -> {a.out}")
