"""Block sparse symmetric-indefinite LDL^T with exact inertia (TEST INFRASTRUCTURE, see oracle/README.md).

IPOPT (behind nlpsol('ipopt'), do_mpc/controller/_mpc.py:1328) factorises the augmented KKT matrix with MUMPS and reads
the inertia off the factorisation.  oracle/ipm.py did that with one dense LAPACK dsytrf of the whole (n+m)^2 matrix, which
is exact but O((n+m)^3): unusable at BASELINE config sizes (cfg3: 15 646 unknowns) and 100x slower than IPOPT on cfg2.
This file is the sparse replacement: a right-looking *block* LDL^T on a partition of the unknowns into dense blocks
(one per scenario-tree node, leaves first), each diagonal block factorised by dsytrf (Bunch-Kaufman), fill handled
dynamically block by block.  By Sylvester's law of inertia  inertia(K) = sum_i inertia(D_i)  where D_i are the successive
Schur-complement diagonal blocks, so the inertia is as exact as with the dense factorisation (tests compare the two).

The partition is only a performance hint: any partition gives a correct factorisation as long as every D_i is
nonsingular; a singular D_i is reported (`ok = False`) and the caller falls back to the dense path.
"""
import numpy as np
import scipy.sparse as sp
from scipy.linalg import lapack


def _inertia_from_sytrf(ldu, ipiv):
    n = ldu.shape[0]
    npos = nneg = nzero = 0
    k = 0
    while k < n:
        if ipiv[k] > 0:
            d = ldu[k, k]
            if d > 0: npos += 1
            elif d < 0: nneg += 1
            else: nzero += 1
            k += 1
        else:
            a, b, c = ldu[k, k], ldu[k + 1, k], ldu[k + 1, k + 1]
            det = a * c - b * b
            tr = a + c
            if det < 0: npos += 1; nneg += 1
            elif det > 0:
                if tr > 0: npos += 2
                else: nneg += 2
            else:
                nzero += 1
                if tr > 0: npos += 1
                elif tr < 0: nneg += 1
                else: nzero += 1
            k += 2
    return npos, nneg, nzero


class BlockPartition:
    """part[i] = block of unknown i; blocks are eliminated in increasing block number."""

    def __init__(self, part):
        part = np.asarray(part, dtype=np.int64)
        self.part = part
        self.nb = int(part.max()) + 1 if part.size else 0
        order = np.argsort(part, kind="stable")
        self.sizes = np.bincount(part, minlength=self.nb)
        self.starts = np.concatenate([[0], np.cumsum(self.sizes)])
        self.idx = [order[self.starts[b]:self.starts[b + 1]] for b in range(self.nb)]
        self.local = np.empty(part.size, dtype=np.int64)
        self.local[order] = np.arange(part.size) - self.starts[part[order]]


class BlockLDL:
    """K = L D L' on the block partition; `inertia`, `ok`, `solve(rhs)` like oracle.ipm._KKT."""

    def __init__(self, K, bp: BlockPartition, pivtol=1e-300):
        self.bp = bp
        self.n = K.shape[0]
        A = sp.coo_matrix(K)
        bi, bj = bp.part[A.row], bp.part[A.col]
        keep = bi >= bj                                   # lower block triangle (row block >= column block)
        bi, bj, li, lj, v = bi[keep], bj[keep], bp.local[A.row[keep]], bp.local[A.col[keep]], A.data[keep]
        key = bj * bp.nb + bi                              # grouped by column block, then row block
        o = np.argsort(key, kind="stable")
        key, bi, bj, li, lj, v = key[o], bi[o], bj[o], li[o], lj[o], v[o]
        cuts = np.flatnonzero(np.diff(key)) + 1
        starts = np.concatenate([[0], cuts]); ends = np.concatenate([cuts, [key.size]])
        sizes = bp.sizes
        # blk[j] = {i: dense K_ij (rows block i, cols block j)} for i >= j
        blk = [dict() for _ in range(bp.nb)]
        for s, e in zip(starts, ends):
            i, j = int(bi[s]), int(bj[s])
            M = np.zeros((sizes[i], sizes[j]))
            np.add.at(M, (li[s:e], lj[s:e]), v[s:e])
            blk[j][i] = M
        self.fac = [None] * bp.nb
        self.Lcol = [None] * bp.nb
        npos = nneg = nzero = 0
        ok = True
        for j in range(bp.nb):
            if sizes[j] == 0:
                self.fac[j] = (np.zeros((0, 0)), np.zeros(0, dtype=np.int32)); self.Lcol[j] = {}
                continue
            col = blk[j]
            D = col.pop(j, None)
            if D is None:
                D = np.zeros((sizes[j], sizes[j]))
            D = np.tril(D) + np.tril(D, -1).T               # symmetric from the lower triangle
            ldu, ipiv, info = lapack.dsytrf(D, lower=1)
            if info != 0:
                ok = False
                p_, n_, z_ = _inertia_from_sytrf(ldu, ipiv) if info > 0 else (0, 0, sizes[j])
                npos += p_; nneg += n_; nzero += max(z_, 1)
                self.fac[j] = None
                break
            p_, n_, z_ = _inertia_from_sytrf(ldu, ipiv)
            npos += p_; nneg += n_; nzero += z_
            if z_ > 0:
                ok = False
                break
            self.fac[j] = (ldu, ipiv)
            rows = sorted(col.keys())
            # X_i = K_ij D^-1  (stored transposed solve: D^-1 K_ij')
            Lc = {}
            for i in rows:
                Xt, info2 = lapack.dsytrs(ldu, ipiv, np.ascontiguousarray(col[i].T), lower=1)
                Lc[i] = (col[i], Xt)                       # K_ij, D^-1 K_ij'
            self.Lcol[j] = Lc
            # Schur update of the trailing blocks: K_ik -= K_ij D^-1 K_kj'   for i >= k in rows
            for a_, k in enumerate(rows):
                Xt_k = Lc[k][1]
                for i in rows[a_:]:
                    upd = Lc[i][0] @ Xt_k                  # (size_i x size_k)
                    tgt = blk[k].get(i)
                    if tgt is None:
                        blk[k][i] = -upd
                    else:
                        tgt -= upd
            blk[j] = None
        self.ok = ok
        self.inertia = (npos, nneg, nzero)

    def solve(self, rhs):
        bp = self.bp
        b = [np.array(rhs[ix], dtype=float) for ix in bp.idx]
        # forward: b_i -= K_ij D_j^-1 b_j
        for j in range(bp.nb):
            if bp.sizes[j] == 0 or not self.Lcol[j]:
                continue
            for i, (Kij, Xt) in self.Lcol[j].items():
                b[i] -= Xt.T @ b[j]
        # diagonal + backward: x_j = D_j^-1 b_j - sum_i (D_j^-1 K_ij') x_i
        x = [None] * bp.nb
        for j in range(bp.nb - 1, -1, -1):
            if bp.sizes[j] == 0:
                x[j] = np.zeros(0); continue
            ldu, ipiv = self.fac[j]
            xj, _ = lapack.dsytrs(ldu, ipiv, b[j], lower=1)
            for i, (Kij, Xt) in self.Lcol[j].items():
                xj = xj - Xt @ x[i]
            x[j] = xj
        out = np.empty(self.n)
        for ix, xv in zip(bp.idx, x):
            out[ix] = xv
        return out
