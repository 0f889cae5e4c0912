"""CPU restatement of the geometry PRE-PROCESSING of the hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference obtains its grids and clouds from Klampt (`geom.convert('VolumeGrid', gridres)` and
`geom.convert('PointCloud', pcres)`, geometryopt.py:58-66: fast marching + Klampt's surface sampler, un-vendored and
absent here).  What is restated here is THIS repository's definition of those conversions: exact
point-triangle distance, side by generalised winding number near the surface and inherited along z elsewhere, vertices +
per-triangle lattices for the cloud.  It is written independently of the product's builders (numpy, straight loops over
triangles and columns; the product's are hostgeom.c / geometry.py on the host and sio_sdf.cu on the device) so that the
device kernels are checked against something that shares no code with them.
"""
import numpy as np


def _pt_tri_dist2(P, a, b, c):
    """Squared distance from every point of P (n,3) to triangle abc (Ericson, Real-Time Collision Detection 5.1.5),
    region by region with masks."""
    ab, ac = b - a, c - a
    ap = P - a
    d1, d2 = ap @ ab, ap @ ac
    out = np.full(len(P), np.nan)
    todo = np.ones(len(P), dtype=bool)

    def settle(mask, val):
        m = todo & mask
        out[m] = val[m]
        todo[m] = False
    settle((d1 <= 0) & (d2 <= 0), (ap * ap).sum(1))
    bp = P - b
    d3, d4 = bp @ ab, bp @ ac
    settle((d3 >= 0) & (d4 <= d3), (bp * bp).sum(1))
    vc = d1 * d4 - d3 * d2
    with np.errstate(divide='ignore', invalid='ignore'):
        v = d1 / (d1 - d3)
        q = ap - (0.0 + v[:, None] * ab)
        settle((vc <= 0) & (d1 >= 0) & (d3 <= 0), (q * q).sum(1))
        cp = P - c
        d5, d6 = cp @ ab, cp @ ac
        settle((d6 >= 0) & (d5 <= d6), (cp * cp).sum(1))
        vb = d5 * d2 - d1 * d6
        w = d2 / (d2 - d6)
        q = ap - (0.0 + w[:, None] * ac)
        settle((vb <= 0) & (d2 >= 0) & (d6 <= 0), (q * q).sum(1))
        va = d3 * d6 - d5 * d4
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        q = bp - (0.0 + w[:, None] * (c - b))
        settle((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), (q * q).sum(1))
        den = 1.0 / (va + vb + vc)
        v, w = vb * den, vc * den
        q = ap - ((0.0 + v[:, None] * ab) + w[:, None] * ac)
        settle(np.ones(len(P), dtype=bool), (q * q).sum(1))
    return out


def _dot3(A, B):
    return A[:, 0] * B[:, 0] + A[:, 1] * B[:, 1] + A[:, 2] * B[:, 2]


def _winding(P, V, T):
    """Generalised winding number of the mesh about every point of P (sum of signed solid angles / 4 pi)."""
    s = np.zeros(len(P))
    for t in T:
        A, B, C = V[t[0]] - P, V[t[1]] - P, V[t[2]] - P
        la, lb, lc = np.sqrt(_dot3(A, A)), np.sqrt(_dot3(B, B)), np.sqrt(_dot3(C, C))
        num = _dot3(A, np.cross(B, C))
        den = la * lb * lc + _dot3(A, B) * lc + _dot3(B, C) * la + _dot3(C, A) * lb
        s = s + 2.0 * np.arctan2(num, den)
    return s / (4.0 * np.pi)


def mesh_sdf(verts, tris, bmin, h, dims):
    """Signed distance at the cell centres bmin + (i+.5, j+.5, k+.5) h, float32 (nx,ny,nz)."""
    V = np.asarray(verts, dtype=np.float64)
    T = np.asarray(tris, dtype=np.int64)
    nx, ny, nz = (int(d) for d in dims)
    ax = [bmin[a] + (np.arange(n) + 0.5) * h[a] for a, n in enumerate((nx, ny, nz))]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    P = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    best = np.full(len(P), np.inf)
    for t in T:
        best = np.minimum(best, _pt_tri_dist2(P, V[t[0]], V[t[1]], V[t[2]]))
    dist = np.sqrt(best).reshape(nx, ny, nz)
    lim = 1.0001 * h[2]
    near = ~((dist[:, :, :-1] > lim) | (dist[:, :, 1:] > lim))          # cell k (>= 1) decides for itself
    need = np.ones((nx, ny, nz), dtype=bool)
    need[:, :, 1:] = near
    side = np.zeros((nx, ny, nz), dtype=np.int8)
    idx = np.nonzero(need.ravel())[0]
    for c0 in range(0, len(idx), 20000):                                   # bounded temporaries
        sel = idx[c0:c0 + 20000]
        side.ravel()[sel] = np.where(_winding(P[sel], V, T) > 0.5, -1, 1)
    out = np.empty((nx, ny, nz), dtype=np.float32)
    s = np.ones((nx, ny), dtype=np.int8)
    for k in range(nz):                                                    # carry the side along every z-column
        s = np.where(side[:, :, k] != 0, side[:, :, k], s)
        out[:, :, k] = (s * dist[:, :, k]).astype(np.float32)
    return out


def surface_sample(verts, tris, res):
    """Vertices + per-triangle lattices (two shorter edges, ceil(edge / res) subdivisions, u + v <= 1, corners left to the
    vertex list), blocks ordered by (m1, m2) then triangle index; points equal after quantisation to 1e-9 m are merged,
    first occurrence kept."""
    V = np.asarray(verts, dtype=np.float64)
    blocks = []
    if np.isfinite(res) and res > 0:
        for t, tri in enumerate(np.asarray(tris, dtype=np.int64)):
            P = V[tri]
            e = [np.sqrt(((P[2] - P[1]) ** 2).sum()), np.sqrt(((P[0] - P[2]) ** 2).sum()), np.sqrt(((P[1] - P[0]) ** 2).sum())]
            apex = int(np.argmax(e))
            a, b, c = P[apex], P[(apex + 1) % 3], P[(apex + 2) % 3]
            m1 = max(1, int(np.ceil(np.sqrt(((b - a) ** 2).sum()) / res - 1e-9)))
            m2 = max(1, int(np.ceil(np.sqrt(((c - a) ** 2).sum()) / res - 1e-9)))
            if m1 == 1 and m2 == 1:
                continue
            pts = []
            for i in range(m1 + 1):
                u = i / m1
                for j in range(m2 + 1):
                    v = j / m2
                    if u + v <= 1.0 + 1e-12 and (i, j) not in ((0, 0), (m1, 0), (0, m2)):
                        pts.append((a * ((1.0 - u) - v) + b * u) + c * v)
            blocks.append(((m1, m2, t), np.array(pts).reshape(-1, 3)))
    blocks.sort(key=lambda kb: kb[0])
    allp = np.vstack([V] + [b for _, b in blocks])
    seen, keep = set(), []
    for i, key in enumerate(map(tuple, np.rint(allp * 1e9).astype(np.int64))):
        if key not in seen:
            seen.add(key)
            keep.append(i)
    return allp[keep]
