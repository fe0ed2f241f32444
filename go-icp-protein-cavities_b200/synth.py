"""Seeded synthetic workloads of the shapes BASELINE.json names; no network, no datasets.

  bo1_pairs(n, seed)   config #5: cavity pairs shaped like cavities_similar_BO1 (N ~ U{165..306} sites grown as a
                       connected blob on a 1.5 A cubic lattice, shipped colour frequencies, c-FPFH rows of four
                       Dirichlet groups x 200), normalised exactly as jly_main.cpp:72-104 does (centre, joint max-norm
                       scale, 6-significant-digit text round trip).
  deep_pair(seed)      config #4: 100k-point target on a bumpy closed surface, 10k-point rotated/translated/noisy source.
"""
import numpy as np

# colour codes (transformation.hpp:36) with the frequencies counted over the 4 shipped cavities (956 points)
_COLOURS = np.array([16741671, 4646984, 7566712, 30894, 8204959, 15219528, 0, 15231913], dtype=np.int32)
_FREQ = np.array([322, 144, 141, 117, 117, 72, 34, 9], dtype=np.float64)
_DIRS = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.int64)


def random_rotation(rng, max_angle=np.pi):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = rng.uniform(0, max_angle)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def _grow(rng, sites, n):
    """grow a connected lattice blob to n sites (sites: list of int triples)"""
    have = set(map(tuple, sites))
    sites = [tuple(s) for s in sites]
    while len(sites) < n:
        base = sites[rng.integers(len(sites))]
        d = _DIRS[rng.integers(6)]
        cand = (base[0] + int(d[0]), base[1] + int(d[1]), base[2] + int(d[2]))
        if cand not in have:
            have.add(cand)
            sites.append(cand)
    return sites


def _fpfh_rows(rng, n):
    groups = [rng.dirichlet(np.full(k, 0.5), size=n) * 200.0 for k in (11, 11, 11, 8)]
    return np.concatenate(groups, axis=1).astype(np.float32)


def _round6(a):
    """writeNormalizedMolCloudFile (transformation.cpp:340) + loadPointCloud (jly_main.cpp:272): %g text round trip"""
    return np.array([float("%g" % v) for v in np.asarray(a, dtype=np.float64).reshape(-1)], dtype=np.float32).reshape(-1, 3)


def _normalise_pair(src, tgt):
    """jly_main.cpp:72-104 (host I/O path of the reference: centre both, divide by the larger max norm, text round trip)"""
    src = src - src.mean(0)
    tgt = tgt - tgt.mean(0)
    scale = max(np.linalg.norm(src, axis=1).max(), np.linalg.norm(tgt, axis=1).max())
    return _round6(src / scale), _round6(tgt / scale)


def bo1_pair(rng, similar=True):
    n_t = int(rng.integers(165, 307))
    t_sites = _grow(rng, [(0, 0, 0)], n_t)
    t_c = rng.choice(_COLOURS, size=n_t, p=_FREQ / _FREQ.sum()).astype(np.int32)
    t_f = _fpfh_rows(rng, n_t)
    if similar:   # source = target blob with 20 % of the sites resampled + a random rigid motion
        keep = rng.permutation(n_t)[: int(0.8 * n_t)]
        n_s = int(np.clip(n_t + rng.integers(-20, 21), 165, 306))
        s_sites = _grow(rng, [t_sites[k] for k in keep], max(n_s, len(keep)))
        n_s = len(s_sites)
        s_c = np.concatenate([t_c[keep], rng.choice(_COLOURS, size=n_s - len(keep), p=_FREQ / _FREQ.sum()).astype(np.int32)])
        s_f = np.concatenate([t_f[keep], _fpfh_rows(rng, n_s - len(keep))])
        perm = rng.permutation(n_s)
        s_sites = [s_sites[k] for k in perm]
        s_c, s_f = s_c[perm], s_f[perm]
    else:
        n_s = int(rng.integers(165, 307))
        s_sites = _grow(rng, [(0, 0, 0)], n_s)
        s_c = rng.choice(_COLOURS, size=n_s, p=_FREQ / _FREQ.sum()).astype(np.int32)
        s_f = _fpfh_rows(rng, n_s)
    tgt = 1.5 * np.array(t_sites, dtype=np.float64) @ random_rotation(rng).T
    src = 1.5 * np.array(s_sites, dtype=np.float64) @ random_rotation(rng).T + rng.uniform(-3, 3, 3)
    d_xyz, m_xyz = _normalise_pair(src, tgt)
    return dict(model_xyz=m_xyz, model_c=t_c, model_fpfh=t_f, data_xyz=d_xyz, data_c=s_c, data_fpfh=s_f, nd=len(d_xyz))


def bo1_pairs(n, seed=4096, similar=True):
    rng = np.random.default_rng(seed)
    return [bo1_pair(rng, similar) for _ in range(n)]


def deep_pair(seed=1234, nm=100000, nd=10000):
    """config #4"""
    rng = np.random.default_rng(seed)
    u = rng.normal(size=(nm, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    centres = rng.normal(size=(8, 3))
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    amp = rng.uniform(-0.3, 0.3, 8)
    r = np.ones(nm)
    for c, a in zip(centres, amp):
        ang = np.arccos(np.clip(u @ c, -1, 1))
        r += a * np.exp(-0.5 * (ang / 0.4) ** 2)
    tgt = u * r[:, None]
    tgt -= tgt.mean(0)
    tgt *= 0.9 / np.linalg.norm(tgt, axis=1).max()
    idx = rng.permutation(nm)[:nd]
    R = random_rotation(rng)
    t = rng.uniform(-0.3, 0.3, 3)
    # the source is the model moved by the INVERSE motion so that (R, t) registers it back
    src = (tgt[idx] - t) @ R + rng.normal(scale=0.005, size=(nd, 3))
    src = src[rng.permutation(nd)]
    return dict(model_xyz=tgt.astype(np.float32), data_xyz=src.astype(np.float32), nd=nd, R_true=R, t_true=t)
