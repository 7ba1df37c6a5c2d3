"""Parameter sweep over the V-cycle (smoothing steps, Chebyshev interval, coarse solve) on the
numpy/scipy prototype of scripts/mg_prototype.py: Krylov iterations x finest-level operator applications.
Usage: python scripts/mg_sweep.py LEVELS _ _ NONSYM

Numpy/scipy prototype of the geometric-multigrid preconditioner on the nested
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
lam = [float(os.environ.get('SAFETY', '1.1')) * lam_max(l, int(os.environ.get('POWER_ITS', '15'))) for l in range(LEVELS + 1)]
print("lambda_max(Dinv K) estimates", lam)


napp = [0]
def chebyshev(l, b, x, its, lo_frac):
    lmax = lam[l]; lmin = lo_frac * lmax
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    if x is None:
        r = b.copy(); x = np.zeros_like(b)
    else:
        r = b - K[l] @ x
        if l == LEVELS: napp[0] += 1
    if its == 0: return x
    d = (Dinv[l] @ r) / theta
    for k in range(its):
        x = x + d
        if k == its - 1: break
        r = r - K[l] @ d
        if l == LEVELS: napp[0] += 1
        rho_new = 1.0 / (2 * sigma - rho)
        d = rho_new * rho * d + (2 * rho_new / delta) * (Dinv[l] @ r)
        rho = rho_new
    return x

BETAS = {1: [1.12500000000000], 2: [1.02387287570313, 1.26408905371085],
         3: [1.00842544782028, 1.08867839208730, 1.33753125909618],
         4: [1.00391310427285, 1.04035811188593, 1.14863498546254, 1.38268869241000]}
KIND = [1]          # 1: first-kind Chebyshev on [lo, 1] lmax; 4: fourth kind (Lottes); 5: fourth kind, optimised weights
_cheb1 = chebyshev
def chebyshev(l, b, x, its, lo_frac):
    if KIND[0] == 1 or l == 0 or its == 0:
        return _cheb1(l, b, x, its, lo_frac)
    lmax = lam[l]
    if x is None:
        r = b.copy(); x = np.zeros_like(b)
    else:
        r = b - K[l] @ x
        if l == LEVELS: napp[0] += 1
    beta = BETAS[its] if KIND[0] == 5 else [1.0] * its
    d = (4.0 / 3.0) / lmax * (Dinv[l] @ r)
    for k in range(1, its + 1):
        x = x + beta[k - 1] * d
        if k == its: break
        r = r - K[l] @ d
        if l == LEVELS: napp[0] += 1
        d = (2 * k - 1) / (2 * k + 3) * d + (8 * k + 4) / (2 * k + 3) / lmax * (Dinv[l] @ r)
    return x

def make_vcycle(pre, post, lo, coarse_it, coarse_lo, gamma=1):
    def vc(l, b):
        if l == 0:
            return chebyshev(0, b, None, coarse_it, coarse_lo)
        pre_l = (pre[0] if l == LEVELS else pre[1]) if isinstance(pre, tuple) else pre
        post_l = (post[0] if l == LEVELS else post[1]) if isinstance(post, tuple) else post
        x = chebyshev(l, b, None, pre_l, lo) if pre_l > 0 else np.zeros_like(b)
        if pre_l > 0:
            r = b - K[l] @ x
            if l == LEVELS: napp[0] += 1
        else:
            r = b
        rc = P[l - 1].T @ r
        xc = vc(l - 1, rc)
        if gamma == 2 and l - 1 > 0:      # W-cycle: second coarse visit on the coarse residual
            xc = xc + vc(l - 1, rc - K[l - 1] @ xc)
        x = x + P[l - 1] @ xc
        x = chebyshev(l, b, x, post_l, lo)
        return x
    return vc

def pcg(A, b, prec, rtol=1e-10, maxit=400):
    x = np.zeros_like(b); r = b.copy(); z = prec(r); p = z.copy(); rz = r @ z
    r0 = np.linalg.norm(r)
    for it in range(1, maxit + 1):
        q = A @ p; napp[0] += 1
        a = rz / (p @ q)
        x += a * p; r -= a * q
        rn = np.linalg.norm(r)
        if rn <= rtol * r0: return x, it
        z = prec(r); rz2 = r @ z
        p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit

L = LEVELS
b = of.body_force(mf.coords, mf.cells, np.full(mf.n_cells, 2200.0), [0, 0, -9.81]) * free[L]
xref = None
print("pre post lo    coarse(it,lo) cyc | its  fine-applies  applies/it")
configs = []
if os.environ.get("SWEEP", "1") == "1":
    for pre, post in ((2, 2), (1, 1), (1, 2), (2, 1), (0, 2), (0, 3), (3, 3), (1, 3)):
        for lo in (0.1, 0.2, 0.3):
            configs.append((pre, post, lo, 20, 0.02, 1))
    configs += [(2, 2, 0.1, 10, 0.05, 1), (2, 2, 0.1, 40, 0.01, 1), (2, 2, 0.1, 20, 0.02, 2), (1, 1, 0.2, 20, 0.02, 2),
                (2, 2, 0.05, 20, 0.02, 1), (1, 1, 0.05, 20, 0.02, 1)]
elif os.environ.get("SWEEP") == "3":       # third sweep: fewer smoothing steps on the finest level only
    for pre, post in (((1, 2), (1, 2)), ((1, 3), (1, 3)), ((1, 2), (1, 3)), ((1, 4), (1, 4)), ((0, 2), (2, 2)), ((1, 2), (2, 2))):
        for cit, clo in ((30, 0.01),):
            configs.append((pre, post, 0.1, cit, clo, 1))
elif os.environ.get("SWEEP") == "6":       # sixth sweep: safety factor on lambda_max (set through SAFETY, see below)
    configs += [(2, 2, 0.1, 30, 0.01, 1)]
elif os.environ.get("SWEEP") == "5":       # fifth sweep: W-cycle with the 30-step coarse solve
    configs += [(2, 2, 0.1, 30, 0.01, 1), (2, 2, 0.1, 30, 0.01, 2), (2, 2, 0.1, 60, 0.005, 2)]
elif os.environ.get("SWEEP") == "4":       # fourth sweep: kind of Chebyshev polynomial
    for kind in (1, 4, 5):
        for pre, post in ((2, 2), (1, 1), (3, 3)):
            configs.append((pre, post, 0.1, 30, 0.01, 10 + kind))
else:       # second sweep: the coarse solve
    for pre, post in ((2, 2), (1, 1)):
        for cit, clo in ((20, 0.02), (30, 0.01), (40, 0.01), (40, 0.005), (60, 0.005), (80, 0.003), (120, 0.001)):
            configs.append((pre, post, 0.1, cit, clo, 1))
for pre, post, lo, cit, clo, gam in configs:
    napp[0] = 0
    KIND[0] = gam - 10 if gam > 10 else 1
    gam = 1 if gam > 10 else gam
    vc = make_vcycle(pre, post, lo, cit, clo, gam)
    t0 = time.time(); x, it = pcg(K[L], b, lambda r: vc(L, r)); dt = time.time() - t0
    if xref is None: xref = x
    err = np.linalg.norm(x - xref) / np.linalg.norm(xref)
    print(f"{str(pre):>6s} {str(post):>6s} {lo:5.2f}  ({cit:2d},{clo:4.2f})   {'W' if gam == 2 else 'V'}{KIND[0]}  | {it:3d}  {napp[0]:6d}  {napp[0]/it:5.2f}   err {err:.1e} ({dt:.0f}s)", flush=True)
