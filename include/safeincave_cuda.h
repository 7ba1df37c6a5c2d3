/* safeincave_cuda.h — C ABI of libsafeincave_cuda.so (hand-written FP64 CUDA for sm_100a).
 *
 * Drop-in boundary for SafeInCave's per-time-step mechanics hot path.  The reference has NO
 * FFI of its own (it is pure Python over DOLFINx/PETSc/torch-CPU); each entry
 * point below replaces the Python method(s) cited next to it.  All citations are relative to
 * the reference checkout (safeincave/...).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; sic_last_error() gives the message
 *     (thread-local).
 *   - the CALLER owns every buffer (the Python host allocates them as torch CUDA tensors and
 *     passes data_ptr()); the library borrows the pointers for the duration of a call and
 *     enqueues work on the caller's cudaStream_t (passed as void*).  No hidden allocation.
 *   - all floating point is IEEE double.  Tensors are symmetric 3x3 stored as 6 Voigt
 *     components in the reference's order [xx, yy, zz, xy, xz, yz] with TENSORIAL shear
 *     (Utils.py:171-227, 251-283).
 *   - C_T and the operator's geometry are TILED (AoSoA): cells are grouped in tiles of SIC_TILE_CELLS = 128 and a
 *     tile's data is one contiguous block, so that the operator can fetch a whole tile with ONE TMA bulk copy:
 *       C_T entry (r,c) of cell i  at  CT[((i/128)*36 + r*6+c)*128 + i%128]          (SIC_CT_INDEX)
 *       geometry of tile t         at  geom_tiles[t]  (sic_geom_tile_t: grad[12][128], vol[128], conn[4][128])
 *     Warp accesses stay coalesced (32 consecutive cells of one row).
 *   - all other per-cell arrays are SoA: component c of cell i lives at a[c*cell_stride + i]
 *     (cell_stride >= n_cells, multiple of 128: whole tiles for the TMA-staged operator).  6x6 tangents are 36 rows, row-major (r*6+c).
 *   - nodal vectors (u, b, x, y ...) are interleaved: dof = 3*node + component, as in the
 *     reference's vector P1 space (MomentumEquation.py:219).
 */
#ifndef SAFEINCAVE_CUDA_H_
#define SAFEINCAVE_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIC_ABI_VERSION 13
#define SIC_MAX_ELEMS 8   /* non-elastic elements per material */
#define SIC_MAX_THERMO 4

#define SIC_TILE_CELLS 128
#define SIC_CT_INDEX(rc, i) ((((size_t)(i) >> 7) * 36 + (size_t)(rc)) * 128 + ((size_t)(i) & 127))

/* geometry of one tile of 128 cells, as the operator kernel stages it in shared memory (15 360 bytes) */
typedef struct {
  double grad[12][SIC_TILE_CELLS];   /* d(phi_a)/dx_j at row 3*a+j */
  double vol[SIC_TILE_CELLS];
  int32_t conn[4][SIC_TILE_CELLS];
} sic_geom_tile_t;

/* non-elastic element kinds (MaterialProps.py classes) */
enum {
  SIC_ELEM_KELVIN = 1,       /* Viscoelastic          :795-885   params: eta, c11, c12, c44 (C1) */
  SIC_ELEM_DISLOCATION = 2,  /* DislocationCreep      :890-961   params: A, Q, n */
  SIC_ELEM_PRESSURE_SOL = 3, /* PressureSolutionCreep :964-1034  params: A, d, Q */
  SIC_ELEM_DESAI = 4,        /* ViscoplasticDesai     :1037-1562 params: mu_1,N_1,a_1,eta,n,beta_1,beta,m,gamma,sigma_t */
  /* SURVEY 8f row 1 (a material that uses none of these runs the very same kernel code as before they existed) */
  SIC_ELEM_MUNSON_DAWSON = 5,  /* MunsonDawsonCreep         :1971-2346 params: A,Q,n,K0,c,m,alpha_w,beta_w,delta,mu */
  SIC_ELEM_MOHR_COULOMB = 6,   /* MohrCoulombViscoplastic   :1565-1746 params: mu_1,N_1,alpha_F,k_F,alpha_Q,sigma_t (derived :1640-1648) */
  SIC_ELEM_MATSUOKA_NAKAI = 7  /* MatsuokaNakaiViscoplastic :1749-1968 params: mu_1,N_1,k_nfc,cohesive_shift,alpha_Q,sigma_t (:1816-1829) */
};

/* rows of the per-cell Desai state block [SIC_DESAI_ROWS][cell_stride] */
enum {
  SIC_DS_ALPHA = 0, SIC_DS_ALPHA0 = 1, SIC_DS_QSI = 2, SIC_DS_QSI_OLD = 3, SIC_DS_FVP = 4,
  SIC_DS_R = 5, SIC_DS_H = 6, SIC_DS_HSMALL = 7, SIC_DS_P = 8 /* ..13 */,
  SIC_DS_ALPHA_K = 14, SIC_DS_Q = 15 /* ..20 */, SIC_DESAI_ROWS = 21
};

/* rows of the per-cell Munson-Dawson state block [SIC_MD_ROWS][cell_stride] (sic_elem_t.desai points at it) */
enum {
  SIC_MD_ZETA = 0, SIC_MD_ZETA_OLD = 1, SIC_MD_F = 2, SIC_MD_ETS = 3, SIC_MD_R = 4, SIC_MD_H = 5, SIC_MD_HSMALL = 6,
  SIC_MD_P = 7 /* ..12 */, SIC_MD_ZETA_K = 13, SIC_MD_Q = 14 /* ..19 */, SIC_MD_ROWS = 20
};
/* Mohr-Coulomb / Matsuoka-Nakai keep one row: the yield function value Fvp (diagnostic the user hooks read) */
enum { SIC_VP_FVP = 0, SIC_VP_ROWS = 1 };

/* flags of sic_post() */
enum {
  SIC_POST_STRAIN = 1,     /* eps = sym grad u            (compute_total_strain, MomentumEquation.py:326-341) */
  SIC_POST_STRESS = 2,     /* sig = CT:(eps - eps_rhs)    (compute_stress :844-866 / compute_elastic_stress :822-842) */
  SIC_POST_INCREMENT = 4,  /* increment_internal_variables (:428-443, MaterialProps.py:1129-1158) */
  SIC_POST_RATES = 8,      /* compute_eps_ne_rate          (:379-395) */
  SIC_POST_ERROR = 16      /* ||eps_k-eps||^2, ||eps||^2   (Simulators.py:433-435) */
};

typedef struct {
  int32_t kind;        /* SIC_ELEM_* */
  int32_t param_off;   /* offset of this element's parameters inside a material-table row */
  double* eps_old;     /* [6][cell_stride]  eps_ne_old       */
  double* rate_old;    /* [6][cell_stride]  eps_ne_rate_old  */
  double* rate;        /* [6][cell_stride]  eps_ne_rate      */
  double* eps_k;       /* [6][cell_stride]  eps_ne_k         */
  double* desai;       /* internal-state block of the element or NULL: Desai [SIC_DESAI_ROWS][cell_stride],
                          Munson-Dawson [SIC_MD_ROWS][cell_stride], Mohr-Coulomb / Matsuoka-Nakai [SIC_VP_ROWS][cell_stride] */
} sic_elem_t;

/* Everything a constitutive / assembly kernel needs.  Filled by the host, passed by pointer. */
typedef struct {
  int32_t abi_version;
  int32_t n_cells, cell_stride, n_nodes;
  /* mesh (P1 tets): connectivity and the constant shape-function gradients per cell */
  const int32_t* conn;   /* [4][cell_stride] node ids */
  const double* grad;    /* [12][cell_stride] d(phi_a)/dx_j at row 3*a+j */
  const double* vol;     /* [cell_stride] */
  const sic_geom_tile_t* geom_tiles;  /* [cell_stride/128] the same conn/grad/vol, tiled for the operator kernel */
  /* scatter plan of the operator: per tile of 128 cells, its unique nodes (tile-interior ones first) and, per
   * unique node, the (cell,slot) references inside the tile.  The kernel stages the 12 nodal forces of every
   * cell in shared memory, sums them per unique node there, and touches global memory once per unique node
   * (plain store for tile-interior nodes, one FP64 atomic otherwise) instead of 12 atomics per cell. */
  const int32_t* tile_ptr;    /* [n_tiles+1] offsets into tile_nodes */
  const int32_t* tile_nint;   /* [n_tiles]   number of tile-interior nodes (listed first) */
  const int32_t* tile_nodes;  /* [tile_ptr[n_tiles]] node ids */
  const int32_t* ent_ptr;     /* [tile_ptr[n_tiles]+1] offsets into ent, per unique (tile,node) */
  const uint16_t* ent;        /* [4*cell_stride] local_cell*4 + slot, grouped by unique (tile,node) */
  /* material: deduplicated parameter rows + per-cell row index */
  const int32_t* mat_id;     /* [cell_stride] */
  const double* mat_table;   /* [n_rows][row_len] */
  int32_t n_rows, row_len;
  int32_t spring_off;        /* row offset of c11,c12,c44 (C) and ci11,ci12,ci44 (C_inv) */
  int32_t n_thermo;          /* Thermoelastic elements (MaterialProps.py:333-382) */
  int32_t thermo_off;        /* row offset of their alpha_th values */
  int32_t n_elems;           /* non-elastic elements */
  sic_elem_t elems[SIC_MAX_ELEMS];
  const double* T;           /* [cell_stride] current temperature  (set_T)  */
  const double* T0;          /* [cell_stride] reference temperature (set_T0) */
  /* per-cell fields */
  double* sig;      /* [6][cell_stride] stress of this iteration          */
  double* sig_k;    /* [6][cell_stride] stress of the previous iteration  */
  double* eps;      /* [6][cell_stride] total strain                      */
  double* eps_prev; /* [6][cell_stride] total strain of previous iteration*/
  double* CT;       /* [cell_stride/128][36][128] consistent tangent, tiled (SIC_CT_INDEX) */
  double* eps_rhs;  /* [6][cell_stride]                                   */
  int32_t* n_singular; /* device counter: cells whose tangent was singular (elastic fallback) */
} sic_problem_t;

const char* sic_last_error(void);
int sic_abi_version(void);
int sic_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- part (1): constitutive update ------------------------------------------------------- */

/* LinearMomentum.compute_CT + compute_eps_rhs (MomentumEquation.py:799-820, 868-890):
 * per cell, G = sum G_i, B = sum B_i (Material.compute_G_B, MaterialProps.py:172-200; FD tangent
 * :640-675; Desai :1432-1562; Kelvin :861-885), CT = inv(C_inv + dt(1-theta)G) (:273-309, singular
 * -> elastic fallback), eps_ne_k_i (:586-605), eps_th (:365-382), eps_rhs.  Reads sig_k. */
int sic_tangent(const sic_problem_t* p, double dt, double theta, void* stream);

/* CT <- C, eps_rhs <- 0: the operator of solve_elastic_response (MomentumEquation.py:892-923). */
int sic_elastic_tangent(const sic_problem_t* p, void* stream);

/* Post-solve phase of one Newton iteration (Simulators.py:416-436), selected by flags.
 *   u            nodal displacement [3*n_nodes] (may be NULL without SIC_POST_STRAIN)
 *   kelvin_phi2  dt(1-theta) of the LAST tangent phase, <0 if none has run yet (then the Kelvin G
 *                is still zero)
 *   err_out      device double[2]: sum (eps_prev-eps)^2 and sum eps^2 over all 9 tensor entries
 *   err_scratch  device double[2*n_blocks] with n_blocks = sic_post_blocks(n_cells)            */
int sic_post(const sic_problem_t* p, const double* u, double dt, double theta, double kelvin_phi2,
             int flags, double* err_out, double* err_scratch, void* stream);
int sic_post_blocks(int n_cells);

/* Commit of a converged step (Simulators.py:509-517): update_internal_variables
 * (MaterialProps.py:1119-1127), update_eps_ne_rate_old (:630-638), update_eps_ne_old (:607-628).
 * The per-element G_i, B_i of the last tangent phase are RE-EVALUATED from sig_k (and the saved
 * Desai alpha_k, Q, P, h, r) instead of being stored (36 doubles per element per cell). */
int sic_commit(const sic_problem_t* p, double dt, double theta, void* stream);

/* update_eps_ne_rate_old alone (Simulators.py:365). */
int sic_commit_rates(const sic_problem_t* p, void* stream);

/* ViscoplasticDesai.compute_initial_hardening (MaterialProps.py:1248-1288) for element `elem`,
 * from the stress in p->sig.  n_clamped (device int) counts alpha_0 <= 1e-6 cells. */
int sic_desai_initial_hardening(const sic_problem_t* p, int elem, double Fvp_0, int32_t* n_clamped,
                                void* stream);

/* ---- multi-GPU: one process per GPU, cells partitioned, interface nodes duplicated ------------- */
#define SIC_MAX_PEERS 16

/* Halo plan of one rank.  Replaces the ghost updates of MomentumEquation.py:915-922, 1018-1025 and the MPI
 * reductions inside PETSc's KSP.  NULL everywhere below means "single GPU". */
typedef struct {
  int32_t n_ranks, rank, n_peers, n_shared_total;
  int32_t peer[SIC_MAX_PEERS];           /* neighbour ranks */
  int32_t peer_off[SIC_MAX_PEERS + 1];   /* [n_peers+1] offsets (in nodes) of each neighbour's slice of idx/buffers */
  const int32_t* idx;      /* [n_shared_total] local node ids shared with each neighbour, same global order on both sides */
  const double* owner_w;   /* [n_nodes] 1.0 where this rank owns the node (lowest rank touching it), else 0.0 */
  double* send_buf;        /* >= 9 * n_shared_total doubles */
  double* recv_buf;        /* >= 9 * n_shared_total doubles */
  void* comm;              /* from sic_comm_init (NCCL path) */
  void* p2p;               /* from sic_p2p_create/connect, or NULL: if set, halo sums and scalar all-reduces are ONE
                              kernel that stores straight into the peers' mailboxes over NVLink (no NCCL call) */
  /* fused operator + halo exchange (multigrid V-cycle, needs p2p; NULL / 0: the exchange is a kernel of its own):
   * the tiles of 128 cells in launch order, those that touch an interface node FIRST.  The operator kernel then carries
   * n_peers * (blocks per peer) extra CTAs right behind the interface tiles which wait for them, exchange the interface
   * sums over NVLink and add what arrives while the interior tiles are still being processed. */
  const int32_t* tile_order;   /* [cell_stride / 128] */
  int32_t n_iface_tiles;
  int32_t reserved;
} sic_halo_t;

/* NCCL communicator over NVLink/NVSwitch (libnccl.so.2 is dlopen'ed on first use; single-GPU runs never touch it).
 * Rank 0 calls sic_comm_unique_id and broadcasts the 128 bytes (e.g. through torch.distributed). */
int sic_comm_unique_id(uint8_t* id128);
int sic_comm_init(const uint8_t* id128, int rank, int n_ranks, void** comm);
int sic_comm_destroy(void* comm);
/* Peer-to-peer exchange over NVLink without NCCL.  Every rank owns a mailbox in its own HBM (cudaMalloc +
 * cudaIpcGetMemHandle); peers map it (cudaIpcOpenMemHandle) and write their interface partial sums and their
 * scalar partial sums directly into it, then raise a flag; the same kernel waits for the peers' flags and adds
 * what arrived.  One launch replaces pack + ncclSend/ncclRecv group + unpack + ncclAllReduce.
 *   sic_p2p_create  : allocate this rank's mailbox (capacity: cap_nodes interface nodes per peer), return its
 *                     64-byte IPC handle in handle64
 *   sic_p2p_connect : all_handles = n_ranks * 64 bytes gathered from every rank (e.g. torch.distributed.all_gather) */
int sic_p2p_create(int rank, int n_ranks, int cap_nodes, void** p2p, uint8_t* handle64);
int sic_p2p_connect(void* p2p, const uint8_t* all_handles);
int sic_p2p_destroy(void* p2p);
int sic_p2p_error(void* p2p);   /* 1 if a wait on a peer's flag timed out (synchronises the device) */
/* halo sum of vec (ncomp values per node; ncomp = 0: none) fused with the sum over ranks of n_scal doubles at
 * `scal` (device, in place; n_scal = 0: none). */
int sic_exchange(const sic_halo_t* h, double* vec, int ncomp, double* scal, int n_scal, void* stream);

/* vec[(node)*ncomp + c] += sum over the other ranks' copies, for every interface node (ncclSend/ncclRecv in one group). */
int sic_halo_sum(const sic_halo_t* h, double* vec, int ncomp, void* stream);
/* in-place sum over ranks of `count` device doubles. */
int sic_allreduce_sum(void* comm, double* dev_buf, int count, void* stream);

/* ---- part (2): matrix-free tangent / RHS "assembly" -------------------------------------- */

/* y = K x (LOCAL cells only; on several GPUs follow with sic_halo_sum) with K = sum_e V_e B^T W CT_e B
 * (a(u,v) of MomentumEquation.py:1008-1011), rows and
 * columns of constrained dofs treated as in assemble_matrix(bcs): fixed[dof] != 0 -> y[dof] = x[dof].
 * x must already be zero on fixed dofs for the symmetric elimination to hold (the solvers do that). */
int sic_apply(const sic_problem_t* p, const double* x, double* y, const uint8_t* fixed, void* stream);

/* r = b_ext - sum_e V_e B^T W CT_e (B x0 - eps_rhs_e), then r[fixed] = 0:
 * linear form of MomentumEquation.py:1014-1020 (b_rhs + body + Neumann) with apply_lifting/set_bc. */
int sic_residual0(const sic_problem_t* p, const double* b_ext, const double* x0, double* r,
                  const uint8_t* fixed, const sic_halo_t* halo, void* stream);

/* nodal 3x3 diagonal blocks of K, inverted, fixed dofs decoupled: dinv[9*node..] (block Jacobi) */
int sic_block_jacobi(const sic_problem_t* p, double* dinv, const uint8_t* fixed, const sic_halo_t* halo, void* stream);

/* Neumann load of MomentumBC.py:247-277: b[3*node+c] += int_F (p_bc + rho_bc g_bc (H_bc - x_dir)) n_c phi ds.
 *   tri      [3][n_tri] node ids;  area_n [3][n_tri] outward normal * area;  bc_of_tri [n_tri] -> bc or -1
 *   bc_par   [n_bc][4] = {p(t) (already negated as in :275), rho*g, H, direction}                     */
int sic_neumann(int n_tri, const int32_t* tri, const double* area_n, const int32_t* bc_of_tri,
                const double* coords, int n_bc, const double* bc_par, double* b, void* stream);

/* ---- part (3): Krylov solve ---------------------------------------------------------------- */
enum {
  SIC_KSP_CG = 1,        /* textbook preconditioned CG: two reductions per iteration */
  SIC_KSP_BICGSTAB = 2,
  SIC_KSP_CGCG = 3       /* Chronopoulos-Gear CG: the same iterates in exact arithmetic with ONE fused reduction per
                            iteration (gamma = r.u, delta = u.Ku, ||r||^2).  SYMMETRIC operators only: measured to diverge
                            on the non-symmetric finite-difference tangent of the creep elements, where
                            textbook CG still converges; opt-in (KSP.single_reduction), never a default */
};

typedef struct {
  int32_t method;         /* SIC_KSP_* */
  int32_t max_it;
  double rtol;            /* on ||r||_2 / ||r0||_2, r0 = b - K x0 restricted to free dofs */
  double atol;
  int32_t check_every;    /* host looks at the residual every this many iterations */
  int32_t use_graph;      /* sic_mg_solve: replay every Krylov iteration (operator, scalar recurrences, the whole V-cycle,
                             exchanges) from ONE captured CUDA graph instead of ~110-170 kernel launches; re-captured
                             whenever a set-up changed the baked-in Chebyshev coefficients.  Ignored by sic_ksp_solve and
                             sic_heat_step (3-5 launches per iteration, dominated by one large operator kernel) */
  int32_t guess_nonzero;  /* x holds an initial guess on the free dofs; rtol is then relative to the residual of the
                             ZERO guess (PETSc's default ||r|| < rtol ||b||), so a warm start saves iterations */
  /* results */
  int32_t iterations;
  int32_t reason;         /* >0 converged (2 rtol, 3 atol), <0 diverged (-3 max_it, -9 nan) as PETSc */
  double rnorm, rnorm0;
  /* measurement: with time_operator != 0 the first operator launch of every batch is bracketed by
   * CUDA events on the solve's stream; op_ms accumulates their durations over op_samples launches */
  int32_t time_operator;
  int32_t op_samples;
  double op_ms;
  /* host-side launch accounting of the Krylov loop: iterations replayed as one graph launch each, iterations
   * launched kernel by kernel (all of them without use_graph; with it, the timed iteration of each batch) */
  int32_t graph_launches;
  int32_t direct_iterations;
  /* sic_mg_solve with time_operator: op_ms / op_samples time the finest-level operator INSIDE the V-cycle (the
   * compressed one when the levels carry pc_ct), op_dot_ms / op_dot_samples the exact Krylov operator (k_mg_ebe_dot) */
  int32_t op_dot_samples;
  int32_t xchg_samples;    /* several GPUs: xchg_ms / xchg_samples time one finest-level halo exchange (3 components) per
                              solve, from the end of the operator kernel to the end of the exchange kernel, i.e. including
                              the wait for the slowest neighbour */
  double op_dot_ms;
  double xchg_ms;
} sic_ksp_t;

/* Workspace: sic_ksp_workspace_doubles(n_nodes, method) doubles, caller-allocated. */
int64_t sic_ksp_workspace_doubles(int n_nodes, int method);

/* Solve K u = b for the free dofs with u = x (in: initial guess incl. prescribed values on fixed
 * dofs; out: solution).  Replaces solver.solve of MomentumEquation.py:1023-1025 / 920-922.
 * dinv: block-Jacobi blocks from sic_block_jacobi. */
int sic_ksp_solve(const sic_problem_t* p, sic_ksp_t* ksp, const double* b_ext, double* x,
                  const uint8_t* fixed, const double* dinv, double* work, const sic_halo_t* halo, void* stream);

/* Initial guess of the next Krylov solve by extrapolating the Newton iterates of the current time step (the
 * reference starts every solve from zero, MomentumEquation.py:1023-1025; only the starting point changes, the solve
 * still runs to rtol).  u0 = latest solution, u1, u2 (, u3) the ones before; n_iterates = 3 or 4.  Writes
 *   x = u0 + a (u0 - u1) + b (u1 - u2),   (a, b) = least-squares fit of (u0 - u1) by (u1 - u2) and (u2 - u3),
 * or a = <d1,d0>/<d0,d0>, b = 0 when only the one-term model explains the last increment; a = b = 0 (x = u0) when
 * neither does, when the prediction is not a contraction, or with three iterates (nothing to validate against).
 * Several GPUs: inner products use the owner weights of `halo` and are summed over the ranks (same a, b everywhere).
 * work: sic_guess_workspace_doubles(n_nodes) doubles.  coef_out (host double[5], may be NULL; synchronises):
 * {a, b, terms used, misfit of the one-term model, misfit of the two-term model}. */
int64_t sic_guess_workspace_doubles(int n_nodes);
int sic_guess_extrapolate(int n_nodes, int n_iterates, const double* u0, const double* u1, const double* u2,
                          const double* u3, double* x, const sic_halo_t* halo, double* work, double* coef_out,
                          void* stream);

/* ---- part (3b): geometric multigrid preconditioner on a nested red-refinement hierarchy ------------- */
/* The synthetic 10M-80M cell meshes of BASELINE config 5 are built by regular (Bey) refinement of a gmsh
 * grid, so every level of the hierarchy is available.  P1 spaces on nested meshes are nested, children of a
 * cell have equal volume and the parent's shape-function gradients, hence the GALERKIN coarse operator
 * P^T K P is exactly the same matrix-free operator evaluated with the arithmetic mean of the eight children's
 * C_T: coarse levels are sic_problem_t's of their own and reuse the operator kernel.  Smoother: Chebyshev
 * polynomial in (block-Jacobi)^-1 K; cycle: V(nu,nu); coarsest level: a longer Chebyshev sweep.  Used as the
 * preconditioner of CG in place of the reference's PETSc PC (MomentumEquation.py:1023-1025; `gamg` in
 * examples/mechanics/nobian/run_interlayer.py:2114-2116).  Level 0 is the COARSEST mesh. */
#define SIC_MG_MAX_LEVELS 8

typedef struct {
  sic_problem_t prob;      /* operator of the level: mesh, tiles, scatter plan, CT (coarse levels: CT is an output of
                              sic_mg_setup; constitutive fields may be NULL) */
  const uint8_t* fixed;    /* [3 n_nodes] Dirichlet mask of the level */
  double* dinv;            /* [9 n_nodes] inverted nodal blocks (output of sic_mg_setup) */
  double lambda_max;       /* largest eigenvalue estimate of dinv*K (output of sic_mg_setup) */
  /* relation to the next COARSER level (NULL on level 0) */
  const int32_t* parent_a; /* [n_nodes] coarse node(s) this node interpolates from: a == b for a node that */
  const int32_t* parent_b; /*           exists on the coarse mesh, else the two ends of the bisected edge  */
  const int32_t* rst_ptr;  /* [n_coarse_nodes + 1] CSR over coarse nodes ...                                 */
  const int32_t* rst_idx;  /* [2 n_nodes] ... of the nodes of THIS level they restrict from, weight 1/2 each */
  const int32_t* children; /* [8][n_coarse_cells] cells of this level that refine each coarse cell           */
  /* work vectors, [3 n_nodes] each */
  double *x, *b, *r, *d, *t;
  double* pv;              /* [3 n_nodes] or NULL: power-iteration vector kept BETWEEN calls of sic_mg_setup (warm start) */
  /* compressed operator of the preconditioner, or both NULL: every operator application INSIDE the V-cycle (smoother,
   * residual, power iteration) then reads the symmetric part of C_T as 21 floats per cell and the gradients + volume
   * as 13 floats -- 152 B per cell instead of 408, 144 with the 8 bytes of pc_lidx instead of the 16 of the connectivity -- with FP64 arithmetic; the Krylov operator stays exact */
  float* pc_ct;            /* [cell_stride/128][21][128] filled by sic_mg_setup from prob.CT */
  const float* pc_geom;    /* [cell_stride/128][13][128] rows 0-11 = prob.grad, row 12 = prob.vol, tiled by 128 cells like
                              pc_ct (filled by the caller) */
  float* pc_dinv;          /* [9 n_nodes] or NULL: float copy of dinv, filled by sic_mg_setup, read by the smoother kernels */
  const uint16_t* pc_lidx; /* [4][cell_stride] position of each cell node in its tile's list of unique nodes
                              (prob.tile_nodes from prob.tile_ptr[tile] on): x is gathered once per unique node of a tile
                              into shared memory, the connectivity is not read (filled by the caller) */
  const sic_halo_t* halo;  /* several GPUs: this level's cells are partitioned exactly as for sic_ksp_solve.  Partitioned
                              levels are the finest one and any run of levels below it (never level 0), NESTED: a cell
                              lives on the rank of its ancestor, so the transfer tables between two partitioned levels are
                              rank-local.  The lowest partitioned level's parent_a/b, rst_* index the (replicated) level
                              below it globally and its `children` holds -1 for cells of other ranks.  NULL: replicated /
                              one GPU */
} sic_mg_level_t;

typedef struct {
  int32_t nu;              /* Chebyshev degree of the pre- and of the post-smoother (2) */
  int32_t coarse_its;      /* Chebyshev steps on level 0 (20) */
  double smooth_lo;        /* smoother targets eigenvalues in [smooth_lo, 1] * lambda_max (0.1) */
  double coarse_lo;        /* same for the coarsest level (0.02) */
  double safety;           /* lambda_max is the power-iteration estimate times this (1.15) */
  int32_t power_its;       /* power iterations per level in sic_mg_setup (>= 2, default 16); 0: keep lambda_max as it is */
  int32_t power_its_warm;  /* passes when restarting from pv of the previous setup (4); 0: always start cold */
  int32_t fused_coarse;    /* != 0: the coarsest level's Chebyshev sweep runs as ONE cooperative launch (k_mg_coarse_fused)
                              instead of two launches per step, whenever its grid fits the device co-resident */
  int32_t reserved;
} sic_mg_opts_t;

/* Once per tangent: restrict C_T down the hierarchy (mean of the 8 children), build the block-Jacobi blocks of
 * every level and estimate lambda_max.  levels[n_levels-1].prob is the fine problem (its CT is the input). */
int64_t sic_mg_workspace_doubles(int n_cells, int n_nodes);   /* of the FINEST level; one workspace serves all calls */
int sic_mg_setup(sic_mg_level_t* levels, int n_levels, const sic_mg_opts_t* opts, double* work, void* stream);

/* CG preconditioned by one V-cycle, same contract as sic_ksp_solve (ksp->method is ignored). */
int sic_mg_solve(sic_mg_level_t* levels, int n_levels, const sic_mg_opts_t* opts, sic_ksp_t* ksp,
                 const double* b_ext, double* x, double* work, void* stream);

/* How many cooperative coarsest-level launches (opts->fused_coarse) this process has made (0: switched off, or the
 * cooperative launch was refused and the launch-per-step sweep is used), and how many times an MG-CG iteration has been
 * captured into a CUDA graph (ksp->use_graph). */
long long sic_mg_fused_coarse_launches(void);
long long sic_mg_graph_captures(void);

/* Several GPUs: the compressed operator inside the V-cycle and the halo exchange of its result run as ONE launch
 * (k_mg_ebe_pc_x) when this switch is on (default OFF: measured neutral to slower than two launches, csrc/mg.cu) and the
 * level's halo plan has P2P mailboxes and a tile order (sic_halo_t.tile_order); the counter says how many such launches
 * this process has made. */
void sic_mg_set_fused_exchange(int on);
long long sic_mg_fused_exchange_launches(void);

/* z = V-cycle(r) alone (tests, and users who bring their own Krylov method): reads levels[top].b, writes .x */
int sic_mg_vcycle(sic_mg_level_t* levels, int n_levels, const sic_mg_opts_t* opts, double* work, void* stream);

/* ---- heat equation of the thermo-mechanical path (HeatEquation.py:304-343) ---------------- */
/* Backward Euler, P1:  (M/dt + K + R) T = (M/dt) T_old + q  with  M = int rho cp phi_a phi_b dx,
 * K = int k grad phi_a . grad phi_b dx, R = sum_robin int h phi_a phi_b ds, q_a = sum_neumann int q phi_a ds +
 * sum_robin int h T_inf phi_a ds (HeatBC.py:283-334), Dirichlet nodes prescribed (HeatBC.py:247-281).  Matrix-free
 * (exact P1 element matrices), Jacobi-preconditioned CG with device-resident scalars.  Replaces HeatDiffusion.solve. */
typedef struct {
  int32_t n_cells, cell_stride, n_nodes, n_tri;
  const int32_t* conn;     /* [4][cell_stride]  as sic_problem_t */
  const double* grad;      /* [12][cell_stride] */
  const double* vol;       /* [cell_stride] */
  const double* rho_cp;    /* [cell_stride] density * specific heat capacity (Material.density, .cp) */
  const double* k;         /* [cell_stride] thermal conductivity (Material.k) */
  const int32_t* tri;      /* [3][n_tri] boundary triangles */
  const double* tri_area;  /* [n_tri] */
  const double* tri_h;     /* [n_tri] Robin coefficient h of the facet's BC (0: none) */
  const double* tri_q;     /* [n_tri] Neumann flux + h * T_inf of the facet's BCs at the current time */
  const uint8_t* fixed;    /* [n_nodes] Dirichlet mask */
  const sic_halo_t* halo;  /* several GPUs: cells and boundary triangles partitioned as for the momentum equation (the
                              reference's heat solve is MPI-parallel, HeatEquation.py:344-364); NULL: one GPU */
} sic_heat_t;

int64_t sic_heat_workspace_doubles(int n_nodes);
/* One backward-Euler step.  T: in = initial guess with the prescribed values on fixed nodes, out = solution.
 * ksp: rtol (relative to the residual of the zero guess, as PETSc's default), atol, max_it, check_every in;
 * iterations, reason, rnorm out. */
int sic_heat_step(const sic_heat_t* h, double dt, const double* T_old, double* T, sic_ksp_t* ksp, double* work,
                  void* stream);
/* get_T_elems (HeatEquation.py:286-302): cell value = mean of the four nodal values */
int sic_heat_cell_mean(const sic_heat_t* h, const double* T_nodes, double* T_cells, void* stream);

/* ---- output fields of a converged step (outside the timed hot path) --------------------------------------- */
/* compute_p_elems / q_elems / p_nodes / q_nodes (MomentumEquation.py:287-324, 944-976) with the smoother of
 * Grid.py:198-242: p = tr(sig)/3, q = sqrt(3 J2) per cell from p->sig; node value = volume-weighted average over the
 * incident cells (grid.A_csr); element value = mean of its four node values (grid.B_csr).
 *   node_vol  [n_nodes] sum of the incident cells' volumes: sic_node_volumes, once per mesh
 *   p_elems / q_elems [n_cells] may be NULL.  Several GPUs: the nodal sums are completed by the halo sum. */
int sic_node_volumes(const sic_problem_t* p, double* node_vol, const sic_halo_t* halo, void* stream);
int sic_pq_fields(const sic_problem_t* p, const double* node_vol, double* p_nodes, double* q_nodes, double* p_elems,
                  double* q_elems, const sic_halo_t* halo, void* stream);

/* ---- measurement helpers --------------------------------------------------------------------- */
/* dependent-free DFMA chains; returns achieved FLOP/s in *flops (used to record the FP64 peak) */
int sic_fp64_peak(double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAFEINCAVE_CUDA_H_ */
