"""TEST INFRASTRUCTURE -- numpy model of the device's neighbour-list build (enhancershoebox_b200/csrc/esb_kernels.cu,
build_lists): the same cell grid, candidate order, per-bead spans (row ranges, slab distances with their slack, x cells
from sqrt(reach^2 - d^2), half-list clip, row ranges clipped to the rows a class occupies, small classes as one span
gated by their bounding box) in float32, so that the GEOMETRY of the build can be checked on the CPU against a brute-force
pair search on configurations the GPU tests cannot enumerate (beads on cell and row boundaries, at the walls, classes of
1 / 32 / 33 beads, empty classes).  Stands in for LAMMPS' binned build (`neighbor 0.3 multi`, run_single_shoebox.py:532):
the list must hold every pair within cut + skin, once.  Not used by the product."""
import numpy as np

F = np.float32
NCLASS = 6


def cell_coord(x, lo, inv, n):
    c = ((F(x) - F(lo)) * F(inv)).astype(np.int64) if isinstance(x, np.ndarray) else int((F(x) - F(lo)) * F(inv))
    return np.clip(c, 0, n - 1) if isinstance(c, np.ndarray) else min(max(c, 0), n - 1)


class Grid:
    def __init__(self, lo, hi, n_pad, cell_edge=2.8, xcell_edge=1.0):
        """rebuild_setups (esb_api.cu): rows of edge cell_edge over (y, z) anchored at lo, ncx cells along x."""
        self.lo, self.hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
        cap = 3 * n_pad - 32
        c = cell_edge
        while True:
            self.nry = max(1, int(np.ceil((hi[1] - lo[1]) / c)))
            self.nrz = max(1, int(np.ceil((hi[2] - lo[2]) / c)))
            if self.nry * self.nrz * NCLASS <= cap and self.nry <= 255 and self.nrz <= 255:
                break
            c *= 1.05
        self.edge = c
        self.ncx = max(1, min(int(np.ceil((hi[0] - lo[0]) / xcell_edge)), cap // (self.nry * self.nrz * NCLASS)))
        self.nrows = self.nry * self.nrz
        self.rinv, self.redge = F(1.0 / c), F(c)
        self.xinv = F(self.ncx / (hi[0] - lo[0]))
        self.lo32 = self.lo.astype(F)


def reach_table(cut, class_of, skin=0.3):
    """fill() in rebuild_setups: reach[type][class] = largest (cut + skin)(1 + 1e-6) + 1e-6 towards any type of the class."""
    nt = cut.shape[0]
    reach = np.zeros((nt, NCLASS), F)
    for a in range(nt):
        for b in range(nt):
            if cut[a, b] > 0:
                reach[a, class_of[b]] = max(reach[a, class_of[b]], F((cut[a, b] + skin) * (1.0 + 1e-6) + 1e-6))
    return reach


def build_half_list(x, types, class_of, cut, grid, skin=0.3):
    """Returns the list of (i, j) the device build would keep (half list: each unordered pair under the bead that comes first
    in candidate order) and the number of distance tests."""
    x = np.asarray(x, F)
    n = len(x)
    cls = class_of[types]
    g = grid
    ry = cell_coord(x[:, 1], g.lo32[1], g.rinv, g.nry)
    rz = cell_coord(x[:, 2], g.lo32[2], g.rinv, g.nrz)
    cx = cell_coord(x[:, 0], g.lo32[0], g.xinv, g.ncx)
    slot = (cls * g.nrows + rz * g.nry + ry) * g.ncx + cx
    nslots = g.nrows * NCLASS * g.ncx
    sidx = np.lexsort((np.arange(n), slot))                        # candidate order: (slot, atom index)
    cend = np.concatenate([[0], np.cumsum(np.bincount(slot, minlength=nslots))])   # cend[s] = start(s), cend[s + 1] = end(s)
    csk = ((cut + skin) ** 2).astype(F)
    reach = reach_table(cut, class_of, skin)
    pop = np.array([cend[(B + 1) * g.nrows * g.ncx] - cend[B * g.nrows * g.ncx] for B in range(NCLASS)])
    crow, cbox = [], []
    for B in range(NCLASS):                                        # rows a class occupies; bounding box of a small class
        m = cls == B
        crow.append((ry[m].min(), ry[m].max(), rz[m].min(), rz[m].max()) if m.any() else (255, 0, 255, 0))
        cbox.append((x[m].min(0), x[m].max(0)) if m.any() else None)
    pairs, tests = [], 0
    for sp_i in range(n):
        i = sidx[sp_i]
        p, ti, A = x[i], types[i], cls[i]
        for B in range(A, NCLASS):
            c = reach[ti, B]
            if not (c > 0 and pop[B] > 0):
                continue
            spans = []
            if pop[B] <= 32:
                lo_b, hi_b = cbox[B]
                e = np.maximum(np.maximum(lo_b - p, p - hi_b), F(0))
                if F(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) < F(c * c):
                    spans.append((cend[B * g.nrows * g.ncx], cend[(B + 1) * g.nrows * g.ncx]))
            else:
                y0 = max(cell_coord(p[1] - c, g.lo32[1], g.rinv, g.nry), crow[B][0]); y1 = min(cell_coord(p[1] + c, g.lo32[1], g.rinv, g.nry), crow[B][1])
                z0 = max(cell_coord(p[2] - c, g.lo32[2], g.rinv, g.nrz), crow[B][2]); z1 = min(cell_coord(p[2] + c, g.lo32[2], g.rinv, g.nrz), crow[B][3])
                if B == A:
                    z0 = max(z0, cell_coord(p[2], g.lo32[2], g.rinv, g.nrz))
                for cz in range(z0, z1 + 1):
                    for cy in range(y0, y1 + 1):
                        ylo = F(F(cy) * g.redge + g.lo32[1]); zlo = F(F(cz) * g.redge + g.lo32[2])
                        dy = max(F(max(F(ylo - p[1]), F(p[1] - F(ylo + g.redge))) - F(1e-5)), F(0))
                        dz = max(F(max(F(zlo - p[2]), F(p[2] - F(zlo + g.redge))) - F(1e-5)), F(0))
                        rr = F(c * c - F(dy * dy + dz * dz))
                        if not rr > 0:
                            continue
                        rx = F(rr * F(1.0 / np.sqrt(rr)) * F(1.00001))
                        x0 = cell_coord(p[0] - rx, g.lo32[0], g.xinv, g.ncx); x1 = cell_coord(p[0] + rx, g.lo32[0], g.xinv, g.ncx)
                        base = (B * g.nrows + cz * g.nry + cy) * g.ncx
                        spans.append((cend[base + x0], cend[base + x1 + 1]))
            for ka, kb in spans:
                ka = min(max(ka, sp_i + 1), kb)                    # nothing at or before the bead itself
                for k in range(ka, kb):
                    j = sidx[k]
                    d = p - x[j]
                    tests += 1
                    if F(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < csk[ti, types[j]]:
                        pairs.append((min(i, j), max(i, j)))
    return pairs, tests


def brute_force(x, types, cut, skin=0.3, band=2e-5):
    """(sure, maybe): unordered pairs within cut + skin by more than `band`, and those within `band` of it."""
    x = np.asarray(x, np.float64)
    d = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
    c = cut[types[:, None], types[None, :]]
    iu = np.triu_indices(len(x), 1)
    r, cc = d[iu], c[iu] + skin
    on = c[iu] > 0
    sure = on & (r < cc - band)
    maybe = on & (np.abs(r - cc) <= band)
    as_set = lambda m: set(zip(iu[0][m].tolist(), iu[1][m].tolist()))
    return as_set(sure), as_set(maybe)
