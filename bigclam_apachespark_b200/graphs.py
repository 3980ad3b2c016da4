"""Graph containers either side of the hot path: compact .npz fixtures of the SNAP graphs the
reference ships (data/*.txt), and synthetic generators for BASELINE.json's configs 3-5.

All of this is one-off host-side integer work (CSR construction); the hot path never sees it.
"""
from __future__ import annotations

import os

import numpy as np

# package data: the SNAP topologies the reference ships, as compact .npz (written by tests/golden/make_fixtures.py);
# BIGCLAM_GRAPH_DIR points somewhere else
FIXTURE_DIR = os.environ.get("BIGCLAM_GRAPH_DIR", os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "graphs"))


def csr_from_undirected(n: int, u: np.ndarray, v: np.ndarray):
    """Simple undirected graph given each edge once (u != v) -> symmetric CSR, lists sorted."""
    a = np.concatenate([u, v]).astype(np.int64)
    b = np.concatenate([v, u]).astype(np.int64)
    order = np.lexsort((b, a))
    a, b = a[order], b[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(a, minlength=n), out=rowptr[1:])
    return rowptr, b.astype(np.int32)


def save_npz_graph(path: str, rowptr: np.ndarray, col: np.ndarray, ids: np.ndarray) -> None:
    """Stores the upper triangle (u < v) of a simple undirected CSR, plus the original ids."""
    n = len(rowptr) - 1
    u = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    m = u < col
    contiguous = bool(np.array_equal(ids, np.arange(ids[0], ids[0] + n)))
    np.savez_compressed(
        path, n=np.int64(n), deg_upper=np.bincount(u[m], minlength=n).astype(np.int32),
        v=col[m].astype(np.int32),
        ids=(np.array([ids[0]], dtype=np.int64) if contiguous else np.diff(ids, prepend=0).astype(np.int64)),
        ids_contiguous=np.bool_(contiguous))


def load_npz_graph(name_or_path: str):
    """Returns (rowptr int64, col int32, ids int64) of a fixture written by save_npz_graph."""
    path = name_or_path if os.path.exists(name_or_path) else os.path.join(FIXTURE_DIR, name_or_path + ".npz")
    z = np.load(path)
    n = int(z["n"])
    u = np.repeat(np.arange(n, dtype=np.int32), z["deg_upper"])
    rowptr, col = csr_from_undirected(n, u, z["v"])
    ids = (np.arange(n, dtype=np.int64) + int(z["ids"][0])) if bool(z["ids_contiguous"]) else np.cumsum(z["ids"])
    return rowptr, col, ids.astype(np.int64)


def have_fixture(name: str) -> bool:
    return os.path.exists(os.path.join(FIXTURE_DIR, name + ".npz"))


def synthetic_F0(n: int, k: int, seed: int = 1234, density: float = 0.05) -> np.ndarray:
    """BASELINE.md synthetic init: F0[u,c] = U[0,1) with probability `density`, else 0 (fp64)."""
    rng = np.random.default_rng(seed)
    F = rng.random((n, k))
    F *= rng.random((n, k)) < density
    return F


def synthetic_F0_csr(n: int, k: int, seed: int = 1234, density: float = 0.05, chunk: int = 1 << 20):
    """Same distribution as synthetic_F0 (each entry non-zero with probability `density`, values U[0,1)) as CSR
    rows (indptr, indices, values), generated in row chunks so that n x k never has to exist densely
    (R-MAT 10M x K=1000).  Not entry-for-entry the same matrix as synthetic_F0(seed)."""
    rng = np.random.default_rng(seed)
    indptr = np.zeros(n + 1, dtype=np.int64)
    idx_parts, val_parts = [], []
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        cnt = rng.binomial(k, density, size=m)
        total = int(cnt.sum())
        # distinct components per row: draw, then sort keys of (row, random) and take the first cnt of a permutation
        row = np.repeat(np.arange(m, dtype=np.int64), cnt)
        comp = np.empty(total, dtype=np.int32)
        # rejection-free: for every row a random permutation prefix via argsort of random keys per chunk of rows
        # would need m x k memory; instead draw with replacement and repair duplicates (rare at low density)
        comp[:] = rng.integers(0, k, size=total)
        key = row * k + comp
        _, first = np.unique(key, return_index=True)
        keep = np.zeros(total, dtype=bool)
        keep[first] = True
        row, comp = row[keep], comp[keep]
        order = np.lexsort((comp, row))
        row, comp = row[order], comp[order]
        idx_parts.append(comp)
        val_parts.append(rng.random(len(comp)))
        indptr[lo + 1:lo + m + 1] = np.bincount(row, minlength=m)
    np.cumsum(indptr, out=indptr)
    return indptr, np.concatenate(idx_parts) if idx_parts else np.zeros(0, np.int32), \
        np.concatenate(val_parts) if val_parts else np.zeros(0)


def synthetic_F0_csr_stratified(n: int, k: int, seed: int = 1234, density: float = 0.05, chunk: int = 1 << 21):
    """CSR rows with exactly round(k * density) non-zeros each, one per stride of 1/density components (a uniformly
    random component inside every stride, values U[0,1)): distinct and ascending by construction, so no sort and no
    de-duplication — the generator for sizes where n x k must never exist densely and the i.i.d. variant
    (synthetic_F0_csr) would spend minutes sorting (R-MAT 10M x K=1000: 5e8 entries).  Same marginals per component
    as synthetic_F0 (each component non-zero with probability `density`); the row lengths are constant instead of
    binomial."""
    rng = np.random.default_rng(seed)
    stride = max(1, int(round(1.0 / density)))
    per_row = max(1, k // stride)
    indptr = np.arange(n + 1, dtype=np.int64) * per_row
    indices = np.empty(n * per_row, dtype=np.int32)
    values = np.empty(n * per_row, dtype=np.float64)
    base = (np.arange(per_row, dtype=np.int32) * stride)[None, :]
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        off = rng.integers(0, stride, size=(m, per_row), dtype=np.int32)
        indices[lo * per_row:(lo + m) * per_row] = np.minimum(base + off, k - 1).reshape(-1)
        values[lo * per_row:(lo + m) * per_row] = rng.random(m * per_row)
    return indptr, indices, values


def rmat_graph_fast(scale_n: int, n_edges: int, a=0.57, b=0.19, c=0.19, seed: int = 42, permute: bool = True):
    """Same R-MAT family as rmat_graph with 16-bit random draws per level and a single key sort (about 10x faster at
    1e8 edges); not edge-for-edge the same graph as rmat_graph(seed)."""
    rng = np.random.default_rng(seed)
    bits = int(np.ceil(np.log2(max(scale_n, 2))))
    ta, tb, tc = int(a * 65536), int((a + b) * 65536), int((a + b + c) * 65536)
    keys = np.zeros(0, dtype=np.int64)
    need = n_edges
    while need > 0:
        m = int(need * 1.25) + 1024
        u = np.zeros(m, dtype=np.int64)
        v = np.zeros(m, dtype=np.int64)
        for _ in range(bits):
            r = rng.integers(0, 65536, size=m, dtype=np.uint16)
            right = ((r >= ta) & (r < tb)) | (r >= tc)              # quadrants b, d: column bit set
            down = r >= tb                                          # quadrants c, d: row bit set
            u = (u << 1) | down
            v = (v << 1) | right
        ok = (u < scale_n) & (v < scale_n) & (u != v)
        lo = np.minimum(u[ok], v[ok])
        hi = np.maximum(u[ok], v[ok])
        keys = np.unique(np.concatenate([keys, lo * scale_n + hi]))
        need = n_edges - len(keys)
    if len(keys) > n_edges:
        keys = keys[rng.permutation(len(keys))[:n_edges]]
    u, v = keys // scale_n, keys % scale_n
    if permute:
        perm = rng.permutation(scale_n)
        u, v = perm[u], perm[v]
    # symmetric CSR by one sort of the directed keys
    a_ = np.concatenate([u, v])
    b_ = np.concatenate([v, u])
    key = a_ * scale_n + b_
    key.sort()
    a_ = key // scale_n
    rowptr = np.zeros(scale_n + 1, dtype=np.int64)
    np.cumsum(np.bincount(a_, minlength=scale_n), out=rowptr[1:])
    return rowptr, (key % scale_n).astype(np.int32)


def rmat_graph_torch(scale_n: int, n_edges: int, a=0.57, b=0.19, c=0.19, seed: int = 42, permute: bool = True, device="cuda"):
    """The same R-MAT family generated with torch on a CUDA device (seconds instead of minutes at 1e8 edges): benchmark
    plumbing only.  Deterministic for a given (seed, torch build, GPU model) — every rank of a multi-GPU run builds the
    identical graph on its own device; not edge-for-edge the NumPy generators' graph."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    bits = int(np.ceil(np.log2(max(scale_n, 2))))
    keys = torch.zeros(0, dtype=torch.int64, device=device)
    need = n_edges
    while need > 0:
        m = int(need * 1.25) + 1024
        u = torch.zeros(m, dtype=torch.int64, device=device)
        v = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(bits):
            r = torch.rand(m, generator=g, device=device, dtype=torch.float32)
            right = ((r >= a) & (r < a + b)) | (r >= a + b + c)
            down = r >= a + b
            u = (u << 1) | down.to(torch.int64)
            v = (v << 1) | right.to(torch.int64)
        ok = (u < scale_n) & (v < scale_n) & (u != v)
        lo = torch.minimum(u[ok], v[ok])
        hi = torch.maximum(u[ok], v[ok])
        keys = torch.unique(torch.cat([keys, lo * scale_n + hi]))
        need = n_edges - keys.numel()
        del u, v, r, ok, lo, hi
    if keys.numel() > n_edges:
        keys = keys[torch.randperm(keys.numel(), generator=g, device=device)[:n_edges]]
    u, v = keys // scale_n, keys % scale_n
    if permute:
        perm = torch.randperm(scale_n, generator=g, device=device)
        u, v = perm[u], perm[v]
    key = torch.cat([u * scale_n + v, v * scale_n + u])
    del u, v, keys
    key, _ = torch.sort(key)
    rows = key // scale_n
    rowptr = torch.zeros(scale_n + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=scale_n), 0)
    col = (key % scale_n).to(torch.int32)
    out = rowptr.cpu().numpy(), col.cpu().numpy()
    del key, rows, rowptr, col
    torch.cuda.empty_cache()
    return out


def synthetic_F0_csr_stratified_torch(n: int, k: int, seed: int = 1234, density: float = 0.05, device="cuda"):
    """synthetic_F0_csr_stratified generated on a CUDA device (benchmark plumbing; same construction, torch's RNG)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    stride = max(1, int(round(1.0 / density)))
    per_row = max(1, k // stride)
    indptr = np.arange(n + 1, dtype=np.int64) * per_row
    indices = np.empty(n * per_row, dtype=np.int32)
    values = np.empty(n * per_row, dtype=np.float64)
    base = (torch.arange(per_row, dtype=torch.int32, device=device) * stride)[None, :]
    chunk = 1 << 21
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        off = torch.randint(0, stride, (m, per_row), generator=g, device=device, dtype=torch.int32)
        indices[lo * per_row:(lo + m) * per_row] = torch.clamp(base + off, max=k - 1).reshape(-1).cpu().numpy()
        values[lo * per_row:(lo + m) * per_row] = torch.rand(m * per_row, generator=g, device=device, dtype=torch.float64).cpu().numpy()
    torch.cuda.empty_cache()
    return indptr, indices, values


def rmat_graph(scale_n: int, n_edges: int, a=0.57, b=0.19, c=0.19, seed: int = 42, permute: bool = True):
    """R-MAT (a,b,c,d) edge generator -> simple undirected CSR over n = scale_n nodes
    (n need not be a power of two: endpoints are drawn in the next power of two and rejected
    when >= n).  Self loops and duplicate pairs are removed, ids randomly permuted."""
    rng = np.random.default_rng(seed)
    bits = int(np.ceil(np.log2(max(scale_n, 2))))
    us, vs = [], []
    need = n_edges
    while need > 0:
        m = int(need * 1.3) + 1024
        u = np.zeros(m, dtype=np.int64)
        v = np.zeros(m, dtype=np.int64)
        for _ in range(bits):
            r = rng.random(m)
            right = (r >= a) & (r < a + b) | (r >= a + b + c)      # quadrants b, d: column bit set
            down = r >= a + b                                       # quadrants c, d: row bit set
            u = (u << 1) | down
            v = (v << 1) | right
        ok = (u < scale_n) & (v < scale_n) & (u != v)
        us.append(u[ok])
        vs.append(v[ok])
        lo = np.minimum(np.concatenate(us), np.concatenate(vs))
        hi = np.maximum(np.concatenate(us), np.concatenate(vs))
        key = np.unique(lo * scale_n + hi)
        us, vs = [key // scale_n], [key % scale_n]
        need = n_edges - len(key)
    u, v = us[0][:n_edges], vs[0][:n_edges]
    if permute:
        perm = rng.permutation(scale_n)
        u, v = perm[u], perm[v]
    return csr_from_undirected(scale_n, u, v)
