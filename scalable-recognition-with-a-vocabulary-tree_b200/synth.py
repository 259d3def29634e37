"""Synthetic "SIFT-shaped" descriptors: D-dimensional, integer valued
0..255 like ezSIFT / ORB output (reference: cbir/descriptors/ezsift.py:70, orb.py:31).

The numpy generators are what the golden vectors under tests/golden/ were made from -- do not
change their draw order.  The torch generators fill device buffers for bench.py.
"""
from __future__ import annotations

import numpy as np


def sift_like(n: int, d: int = 128, n_centres: int = 1000, seed: int = 1234,
              scale: float = 30.0, noise: float = 12.0) -> np.ndarray:
    """The survey's generator: mixture of ``n_centres`` exponential centres + gaussian noise,
    rounded and clipped to bytes.  Returns uint8 [n, d]."""
    rng = np.random.default_rng(seed)
    centres = rng.exponential(scale, (n_centres, d))
    idx = rng.integers(0, n_centres, n)
    x = centres[idx] + rng.normal(0.0, noise, (n, d))
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def image_set(n_images: int, n_desc: int, d: int = 128, n_centres: int = 1000,
              centres_per_image: int = 24, seed: int = 1234, scale: float = 30.0,
              noise: float = 12.0, ragged: bool = False):
    """A Holidays-shaped synthetic dataset: every image draws its descriptors from its own
    subset of the mixture so that retrieval has structure.  Returns a list of uint8 [n_i, d]
    arrays (n_i == n_desc unless ``ragged``)."""
    rng = np.random.default_rng(seed)
    centres = rng.exponential(scale, (n_centres, d))
    out = []
    for _ in range(n_images):
        mine = rng.integers(0, n_centres, centres_per_image)
        n_i = int(rng.integers(max(1, n_desc // 2), n_desc + 1)) if ragged else n_desc
        idx = mine[rng.integers(0, centres_per_image, n_i)]
        x = centres[idx] + rng.normal(0.0, noise, (n_i, d))
        out.append(np.clip(np.rint(x), 0, 255).astype(np.uint8))
    return out


def hierarchical_tree(k: int, depth: int, d: int = 128, seed: int = 7, device="cpu"):
    """Random k-ary centroid hierarchy in heap order (children of n: n*k+1..n*k+k) used by
    bench.py when a million-leaf tree is needed without running the build: child = parent +
    noise whose scale shrinks by ``shrink`` per level.  Returns float32 [n_heap, d] (torch)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    n_heap = sum(k ** l for l in range(depth + 1))
    cent = torch.empty((n_heap, d), dtype=torch.float32, device=device)
    cent[0] = 40.0
    off, sigma = 0, 28.0
    for level in range(depth):
        n_l = k ** level
        parents = cent[off:off + n_l]
        child = parents.repeat_interleave(k, dim=0)
        child = child + sigma * torch.randn(child.shape, generator=g, device=device, dtype=torch.float32)
        cent[off + n_l: off + n_l + n_l * k] = child.clamp_(0.0, 255.0)
        off += n_l
        sigma *= 0.62
    return cent


def descriptors_from_tree(cent, k: int, depth: int, n: int, seed: int = 11, noise: float = 4.0,
                          dtype="u8"):
    """n descriptors = a uniformly drawn deepest-level centroid + gaussian noise, rounded and
    clipped to bytes (torch, on ``cent``'s device)."""
    import torch
    dev = cent.device
    g = torch.Generator(device=dev).manual_seed(seed)
    n_leaf = k ** depth
    leaf_off = sum(k ** l for l in range(depth))
    out = torch.empty((n, cent.shape[1]), dtype=torch.uint8 if dtype == "u8" else torch.float32, device=dev)
    step = 1 << 20
    for s in range(0, n, step):
        e = min(n, s + step)
        idx = torch.randint(0, n_leaf, (e - s,), generator=g, device=dev) + leaf_off
        x = cent[idx] + noise * torch.randn((e - s, cent.shape[1]), generator=g, device=dev)
        x = x.round_().clamp_(0.0, 255.0)
        out[s:e] = x.to(out.dtype)
    return out
