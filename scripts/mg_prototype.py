"""Numpy/scipy prototype of the geometric-multigrid preconditioner on the nested
red-refinement hierarchy of cavern_regular: measures PCG iteration counts of
  (a) nodal block-Jacobi (what csrc/solver.cu does today) and
  (b) a V(nu,nu) cycle with Chebyshev/block-Jacobi smoothing, Galerkin-equivalent coarse tangents
      (mean of the 8 children's C_T), Chebyshev coarse solve.
CPU only, test/experiment infrastructure (uses oracle/fem.py)."""
import os, sys, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fem as of
from oracle import constitutive as oc
from safeincave_b200.mesh import TetMesh, red_refine

LEVELS = int(sys.argv[1]) if len(sys.argv) > 1 else 2
NU = int(sys.argv[2]) if len(sys.argv) > 2 else 2
COARSE_IT = int(sys.argv[3]) if len(sys.argv) > 3 else 20
NONSYM = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0

m0 = TetMesh.load_npz(os.path.join(ROOT, "tests/golden/mesh_cavern_regular.npz"))
meshes = [m0]
for l in range(LEVELS):
    meshes.append(red_refine(meshes[-1]))
print("cells per level", [m.n_cells for m in meshes], "nodes", [m.n_nodes for m in meshes])

def prolongation(mc, mf):
    """P1 interpolation coarse->fine for red refinement: coarse nodes keep ids, midpoints follow."""
    Mc, Mf = mc.n_nodes, mf.n_nodes
    # recover the edge list exactly as red_refine does
    cells = mc.cells
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    keys = np.concatenate([np.minimum(cells[:, a], cells[:, b]) * Mc + np.maximum(cells[:, a], cells[:, b]) for a, b in pairs])
    uniq = np.unique(keys)
    lo, hi = uniq // Mc, uniq % Mc
    assert Mc + uniq.size == Mf
    rows = np.concatenate([np.arange(Mc), Mc + np.arange(uniq.size), Mc + np.arange(uniq.size)])
    cols = np.concatenate([np.arange(Mc), lo, hi])
    vals = np.concatenate([np.ones(Mc), 0.5 * np.ones(uniq.size), 0.5 * np.ones(uniq.size)])
    Pn = sp.csr_matrix((vals, (rows, cols)), shape=(Mf, Mc))
    return sp.kron(Pn, sp.identity(3), format="csr")

def fixed_dofs(m):
    names = m.names[2]
    d = []
    for nm, comp in (("West", 0), ("South", 1), ("Bottom", 2)):
        d.append(of.dirichlet_dofs(m.tris, m.tri_tags, names[nm], comp))
    return np.unique(np.concatenate(d))

E, nu = 102e9, 0.3
def tangent(m, seed=0):
    N = m.n_cells
    C = oc.iso_matrix(np.full(N, E), np.full(N, nu))
    return C

# fine tangent: elastic, optionally softened near the cavern + non-symmetric perturbation
mf = meshes[-1]
CT = [None] * (LEVELS + 1)
CTf = tangent(mf)
if NONSYM > 0:
    rng = np.random.default_rng(0)
    cen = mf.coords[mf.cells].mean(axis=1)
    soft = 1.0 / (1.0 + 5.0 * np.exp(-((cen[:, 0]) ** 2 + (cen[:, 1]) ** 2) / 100.0 ** 2))   # softer near the axis
    CTf = CTf * soft[:, None, None]
    CTf = CTf * (1.0 + NONSYM * rng.standard_normal((mf.n_cells, 6, 6)))
CT[LEVELS] = CTf
for l in range(LEVELS, 0, -1):           # children of parent p are 8p..8p+7 (red_refine)
    CT[l - 1] = CT[l].reshape(-1, 8, 6, 6).mean(axis=1)

t0 = time.time()
K, free, Dinv, P = [], [], [], []
for l, m in enumerate(meshes):
    Kl = of.assemble_K(m.coords, m.cells, CT[l])
    fx = fixed_dofs(m)
    fr = np.ones(Kl.shape[0], dtype=bool); fr[fx] = False
    # eliminate fixed rows/cols, unit diagonal
    D = sp.diags(fr.astype(float))
    Kl = (D @ Kl @ D + sp.diags((~fr).astype(float))).tocsr()
    K.append(Kl); free.append(fr)
    # nodal 3x3 block-Jacobi
    M = m.n_nodes
    blk = np.zeros((M, 3, 3))
    Kc = Kl.tocoo()
    sel = (Kc.row // 3) == (Kc.col // 3)
    np.add.at(blk, (Kc.row[sel] // 3, Kc.row[sel] % 3, Kc.col[sel] % 3), Kc.data[sel])
    bi = np.linalg.inv(blk)
    r = (3 * np.arange(M)[:, None, None] + np.arange(3)[None, :, None]) + 0 * np.arange(3)[None, None, :]
    c = (3 * np.arange(M)[:, None, None] + np.arange(3)[None, None, :]) + 0 * np.arange(3)[None, :, None]
    Dinv.append(sp.csr_matrix((bi.ravel(), (r.ravel(), c.ravel())), shape=Kl.shape))
for l in range(LEVELS):
    Pl = prolongation(meshes[l], meshes[l + 1])
    Pl = sp.diags(free[l + 1].astype(float)) @ Pl @ sp.diags(free[l].astype(float))
    P.append(Pl.tocsr())
print(f"setup {time.time()-t0:.1f}s")
# Galerkin check on the first pair (free dofs)
if LEVELS >= 1:
    G = (P[0].T @ K[1] @ P[0]).tocsr()
    Df = sp.diags(free[0].astype(float))
    diff = (G - Df @ K[0] @ Df)
    print("Galerkin: |P^T K1 P - K0| / |K0| =", abs(diff).max() / abs(K[0]).max())

def lam_max(l, its=15):
    rng = np.random.default_rng(1)
    v = rng.standard_normal(K[l].shape[0]) * free[l]
    for _ in range(its):
        w = Dinv[l] @ (K[l] @ v)
        lam = np.linalg.norm(w) / np.linalg.norm(v)
        v = w / np.linalg.norm(w)
    return lam
lam = [1.1 * lam_max(l) for l in range(LEVELS + 1)]
print("lambda_max(Dinv K) estimates", lam)

def chebyshev(l, b, x, its, lo_frac=0.1):
    """Chebyshev iteration on Dinv K for eigenvalues in [lo_frac*lmax, lmax] (hypre/PETSc style)."""
    lmax = lam[l]; lmin = lo_frac * lmax
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    r = b - K[l] @ x if x is not None else b.copy()
    if x is None: x = np.zeros_like(b)
    d = (Dinv[l] @ r) / theta
    for k in range(its):
        x = x + d
        if k == its - 1: break
        r = r - K[l] @ d
        rho_new = 1.0 / (2 * sigma - rho)
        d = rho_new * rho * d + (2 * rho_new / delta) * (Dinv[l] @ r)
        rho = rho_new
    return x

napply = [0] * (LEVELS + 1)
def vcycle(l, b):
    if l == 0:
        return chebyshev(0, b, None, COARSE_IT, lo_frac=0.02)
    x = chebyshev(l, b, None, NU)
    r = b - K[l] @ x
    xc = vcycle(l - 1, P[l - 1].T @ r)
    x = x + P[l - 1] @ xc
    x = chebyshev(l, b, x, NU)
    return x

def pcg(A, b, prec, rtol=1e-12, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = prec(r); p = z.copy(); rz = r @ z
    r0 = np.linalg.norm(r)
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        rn = np.linalg.norm(r)
        if rn <= rtol * r0: return x, it
        z = prec(r); rz2 = r @ z
        p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit

rng = np.random.default_rng(2)
L = LEVELS
# a realistic rhs: gravity body force
b = of.body_force(mf.coords, mf.cells, np.full(mf.n_cells, 2200.0), [0, 0, -9.81]) * free[L]
t0 = time.time(); xj, itj = pcg(K[L], b, lambda r: Dinv[L] @ r); tj = time.time() - t0
print(f"block-Jacobi PCG: {itj} iterations ({tj:.1f}s)")
t0 = time.time(); xm, itm = pcg(K[L], b, lambda r: vcycle(L, r)); tm = time.time() - t0
print(f"MG V({NU},{NU}) coarse {COARSE_IT}: {itm} iterations ({tm:.1f}s); |xm-xj|/|xj| = {np.linalg.norm(xm-xj)/np.linalg.norm(xj):.2e}")
# cost model in fine-operator-apply equivalents: per V-cycle level l does (2*NU - 1) + 1 applies (pre: NU-1, residual 1, post NU)
cost = sum((2 * NU) * 8.0 ** (l - L) for l in range(1, L + 1)) + COARSE_IT * 8.0 ** (-L)
print(f"apply-equivalents per MG iteration: {1 + cost:.2f} -> total {itm*(1+cost):.0f} vs Jacobi {itj}")
