/* TEST INFRASTRUCTURE -- the parity oracle.  Never linked into, imported by or
 * called from the product (ogl_b200/, include/).  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * A single-threaded CPU restatement of the reference's linear-solve hot path
 * (hpsim/OGL 0.5.4).  Two halves:
 *
 *  assembly.cpp  restates code that IS in /root/reference:
 *      HostMatrix/HostMatrixFreeFunctions.C:21-201, HostMatrix/HostMatrix.C:158-732.
 *      Pinned against the golden vectors of unitTests/test_HostMatrix.C:8-107
 *      and against oracle/_ref (the reference's own free functions compiled from
 *      source) on random meshes.
 *
 *  krylov.cpp    restates the arithmetic that lives in the third-party
 *      dependency Ginkgo (github.com/ginkgo-project/ginkgo @
 *      fc86d48b78cebd2b2c5833a2dcf0fe40f615cf19, fetched by
 *      third_party/ginkgo/CMakeLists.txt:6-24; NOT under /root/reference) in the
 *      order of its `reference` executor, driven by OGL's own stopping criterion
 *      StoppingCriterion/StoppingCriterion.C:11-151.  PARITY UNPINNED for the
 *      solve: the reference holds no test pinning iterations, residuals or
 *      solutions; only the criterion restates in-tree code.
 *
 * label = int32, scalar = FP64 (integration-tests.yml:13-14).
 */
#ifndef OGL_ORACLE_H
#define OGL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t orc_label;
typedef double orc_scalar;

/* ---- assembly (assembly.cpp) ------------------------------------------- */

/* HostMatrixFreeFunctions.C:105-201.  upper = upperAddr (col of an upper entry),
 * lower = lowerAddr (row of an upper entry).  Outputs have nrows+2*upper_nnz
 * entries.  upper_nnz == 0 is accepted here (the reference dereferences
 * tmp_lower[0], HostMatrixFreeFunctions.C:157-158: undefined). */
void orc_init_local_sparsity(orc_label nrows, orc_label upper_nnz,
                             int is_symmetric, const orc_label *upper,
                             const orc_label *lower, orc_label *rows,
                             orc_label *cols, orc_label *permute);

/* HostMatrix.C:504-586: merge n_iface local (cyclic) couplings
 * (iface_rows[i], iface_cols[i]), i = running interface index, into the
 * row-major pattern of local_nnz = nrows+2*upper_nnz entries held in the first
 * local_nnz slots of rows/cols/permute (arrays sized local_nnz+n_iface). */
void orc_merge_local_interfaces(orc_label nrows, orc_label upper_nnz,
                                int is_symmetric, orc_label n_iface,
                                const orc_label *iface_rows,
                                const orc_label *iface_cols, orc_label *rows,
                                orc_label *cols, orc_label *permute);

/* HostMatrixFreeFunctions.C:21-102, with the documented intent for `scale`
 * (the reference's symmetric_update loses it to operator precedence, :27-28;
 * see orc_symmetric_update_as_written). */
void orc_symmetric_update(orc_label total_nnz, orc_label upper_nnz,
                          const orc_label *permute, orc_scalar scale,
                          const orc_scalar *diag, const orc_scalar *upper,
                          orc_scalar *out);
void orc_symmetric_update_as_written(orc_label total_nnz, orc_label upper_nnz,
                                     const orc_label *permute, orc_scalar scale,
                                     const orc_scalar *diag,
                                     const orc_scalar *upper, orc_scalar *out);
void orc_non_symmetric_update(orc_label total_nnz, orc_label upper_nnz,
                              const orc_label *permute, orc_scalar scale,
                              const orc_scalar *diag, const orc_scalar *upper,
                              const orc_scalar *lower, orc_scalar *out);
void orc_symmetric_update_w_interface(orc_label total_nnz, orc_label diag_nnz,
                                      orc_label upper_nnz,
                                      const orc_label *permute, orc_scalar scale,
                                      const orc_scalar *diag,
                                      const orc_scalar *upper,
                                      const orc_scalar *iface, orc_scalar *out);
void orc_non_symmetric_update_w_interface(
    orc_label total_nnz, orc_label diag_nnz, orc_label upper_nnz,
    const orc_label *permute, orc_scalar scale, const orc_scalar *diag,
    const orc_scalar *upper, const orc_scalar *lower, const orc_scalar *iface,
    orc_scalar *out);

/* HostMatrix.C:634-704 (the default device path): staging layout
 * [upper(F) | lower(F, asym only) | diag(n) | local iface(n_iface)] followed by
 * a row_gather through ldu_mapping.  No scaling on this path in the reference. */
void orc_gather_from_staging(orc_label total_nnz, const orc_label *permute,
                             const orc_scalar *staging, orc_scalar *out);

/* HostMatrix.C:180-207: concatenation (done by the caller) times -1. */
void orc_negate(orc_label n, const orc_scalar *in, orc_scalar *out);

/* HostMatrix.C:251-306.  Processor interfaces in interface order: neighbour
 * rank nbr[i], size sz[i], faceCells concatenated in face_cells.  Outputs:
 * *n_targets, target_ids/target_sizes (capacity n_proc_ifaces) ascending by
 * rank, send_idxs = per-target concatenation (capacity sum sz). */
void orc_comm_pattern(orc_label n_proc_ifaces, const orc_label *nbr,
                      const orc_label *sz, const orc_label *face_cells,
                      orc_label *n_targets, orc_label *target_ids,
                      orc_label *target_sizes, orc_label *send_idxs);

/* HostMatrix.C:412-466.  face_cells = concatenated faceCells of all processor
 * interfaces in interface order (n_halo entries).  Sort by row only; ties are
 * resolved by running index (a stable sort -- the reference's std::sort leaves
 * tie order unspecified). */
void orc_non_local_pattern(orc_label n_halo, const orc_label *face_cells,
                           orc_label *rows, orc_label *cols, orc_label *permute);

/* HostMatrix.C:708-732: out[k] = neg_coeffs[permute[k]]. */
void orc_non_local_update(orc_label n_halo, const orc_label *permute,
                          const orc_scalar *neg_coeffs, orc_scalar *out);

/* ---- distributed system + Krylov (krylov.cpp) --------------------------- */

/* One rank's share of the row-block distributed system, as OGL hands it to
 * Ginkgo (CsrMatrixWrapper.H:163-210, Partition.H:57-70). */
typedef struct orc_rank_system {
    orc_label n;              /* local rows                                  */
    orc_label nnz;            /* local entries, row-major sorted COO         */
    const orc_label *rows;    /* [nnz]                                       */
    const orc_label *cols;    /* [nnz]                                       */
    const orc_scalar *vals;   /* [nnz]                                       */
    orc_label n_halo;         /* non-local entries == recv buffer length     */
    const orc_label *nl_rows; /* [n_halo] ascending                          */
    const orc_label *nl_cols; /* [n_halo] recv-buffer slot                   */
    const orc_scalar *nl_vals;/* [n_halo]                                    */
    orc_label n_targets;
    const orc_label *target_ids;   /* [n_targets] ascending neighbour ranks  */
    const orc_label *target_sizes; /* [n_targets]                            */
    const orc_label *send_idxs;    /* [sum target_sizes] blocked by target   */
    const orc_scalar *b;      /* [n]                                         */
    orc_scalar *x;            /* [n] in: initial guess, out: solution        */
} orc_rank_system;

enum { ORC_CG = 0, ORC_BICGSTAB = 1, ORC_GMRES = 2 };
/* ISAI: Ginkgo preconditioner::Isai<isai_type::spd> (Preconditioner.H:225-242), GISAI: <general>
 * (:243-260); sparsityPower 1 */
enum { ORC_PRECOND_NONE = 0, ORC_PRECOND_BJ = 1, ORC_PRECOND_ISAI = 2, ORC_PRECOND_GISAI = 3,
       /* exact ILU(0) / IC(0) with exact triangular solves (Preconditioner.H:106-124, 177-196) and
        * ILU(0) with 5 Jacobi-Richardson sweeps per factor (:143-176); trifactor.hpp */
       ORC_PRECOND_ILU = 4, ORC_PRECOND_IC = 5, ORC_PRECOND_IRILU = 6,
       /* Multigrid: PGM aggregation, V cycle, Jacobi-Richardson smoother, CG coarsest solver
        * (Preconditioner.H:261-341); multigrid.hpp */
       ORC_PRECOND_MULTIGRID = 7 };

typedef struct orc_solve_params {
    int solver;               /* ORC_CG ...                                  */
    int precond;              /* ORC_PRECOND_*                               */
    orc_label max_block_size; /* BJ maxBlockSize (Preconditioner.H:93-94)    */
    orc_scalar tolerance;     /* StoppingCriterion.H:167                     */
    orc_scalar rel_tol;       /* :168                                        */
    orc_label min_iter;       /* effective minIter (after adaptation)        */
    orc_label max_iter;       /* already doubled for BiCGStab (:188)         */
    orc_label frequency;      /* effective evaluation frequency              */
    orc_label krylov_dim;     /* GMRES restart length (Ginkgo default 100)   */
    orc_label threads;        /* 1 = reference executor order; >1 = OpenMP   */
    orc_label mg_max_levels;      /* Multigrid maxLevels (<= 0: 9)           */
    orc_label mg_min_coarse_rows; /* minCoarseRows (<= 0: 10)                */
    orc_label mg_coarse_iters;    /* coarseSolverIters (<= 0: 4)             */
} orc_solve_params;

typedef struct orc_solve_result {
    orc_scalar init_residual; /* normalised L1, StoppingCriterion.C:110      */
    orc_scalar final_residual;/* :119                                        */
    orc_label criterion_calls;/* iter_ counter, :79,85,143                   */
    orc_label n_iterations;   /* what OGL reports (BiCGStab: calls/2)        */
    orc_scalar norm_factor;   /* :32-69                                      */
    orc_label n_history;      /* residual history entries written            */
    double seconds;           /* wall time of the iteration loop             */
} orc_solve_result;

/* y = A x for the distributed system (local apply, then += non-local apply on
 * the exchanged halo), rank by rank.  xs/ys: one pointer per rank. */
void orc_dist_spmv(int n_ranks, const orc_rank_system *ranks,
                   const orc_scalar *const *xs, orc_scalar *const *ys);

/* Solve; returns 0 on success.  history (may be NULL, zero-initialised by the
 * caller) receives the normalised residual indexed by criterion call
 * (StoppingCriterion.C:115-117); skipped calls stay 0. */
int orc_solve(int n_ranks, const orc_rank_system *ranks,
              const orc_solve_params *params, orc_solve_result *result,
              orc_scalar *history, orc_label history_cap);

/* Block-Jacobi pieces exposed for direct parity tests: block pointers found by
 * Ginkgo's find_blocks (natural blocks + greedy agglomeration) and the
 * inverted diagonal blocks (row-major, concatenated). */
orc_label orc_bj_find_blocks(orc_label n, const orc_label *row_ptrs,
                             const orc_label *cols, orc_label max_block_size,
                             orc_label *block_ptrs /* [n+1] */);
/* ISAI values over the CSR pattern (w; spd: also wt = transpose), 0 on success */
int orc_isai_generate(orc_label n, const orc_label *row_ptrs, const orc_label *cols,
                      const orc_scalar *vals, int spd, orc_scalar *w, orc_scalar *wt);
void orc_bj_invert_blocks(orc_label n, const orc_label *row_ptrs,
                          const orc_label *cols, const orc_scalar *vals,
                          orc_label n_blocks, const orc_label *block_ptrs,
                          orc_scalar *inv /* sum b^2 */);
/* Multigrid hierarchy of one (local) matrix, exposed level by level for parity tests: level l holds
 * its matrix (CSR) and, unless it is the coarsest, the aggregate of every row (fine -> coarse) */
void *orc_mg_create(orc_label n, const orc_label *row_ptrs, const orc_label *cols, const orc_scalar *vals,
                    int max_levels, orc_label min_coarse_rows, int coarse_iters);
void orc_mg_destroy(void *h);
int orc_mg_levels(const void *h);
int orc_mg_level_info(const void *h, int level, orc_label *n, orc_label *nnz, orc_label *n_coarse);
int orc_mg_level_get(const void *h, int level, orc_label *row_ptrs, orc_label *cols, orc_scalar *vals,
                     orc_label *agg /* NULL on the coarsest level */);
void orc_mg_apply(const void *h, const orc_scalar *r, orc_scalar *z);   /* one V cycle from a zero guess */

/* ILU(0) / IC(0) factors over the CSR pattern of A (strictly lower part = L, upper part = U incl.
 * the diagonal; IC: lower part incl. the diagonal = L, upper part = its transpose); 0 on success,
 * 1 row without diagonal or with a repeated column, 2 IC on a structurally unsymmetric pattern */
int orc_trifactor(int kind, orc_label n, const orc_label *row_ptrs, const orc_label *cols,
                  const orc_scalar *vals, orc_scalar *factors /* [nnz] */);
/* z = M^-1 r with those factors: exact triangular solves (ILU, IC) or 5 + 5 Jacobi-Richardson
 * sweeps (IRILU) */
int orc_trifactor_apply(int kind, orc_label n, const orc_label *row_ptrs, const orc_label *cols,
                        const orc_scalar *factors, const orc_scalar *r, orc_scalar *z);

/* CPU baseline helpers (bench.py): repeat SpMV / run fixed PCG iterations with
 * `threads` OpenMP threads, return seconds. */
double orc_time_spmv(orc_label n, const orc_label *row_ptrs,
                     const orc_label *cols, const orc_scalar *vals,
                     const orc_scalar *x, orc_scalar *y, int reps, int threads);

/* OpenFOAM-native-equivalent CPU baseline (foam_pcg.cpp): face-based lduMatrix::Amul +
 * PCG with the diagonal preconditioner and OpenFOAM's own normFactor / convergence test,
 * on one rank's lduMatrix (cyclic interfaces as rows/cols/bouCoeffs; lower == NULL when
 * symmetric).  psi: in initial guess, out solution.  history[k] = residual after k iterations. */
int orc_foam_pcg(orc_label n, orc_label n_faces, const orc_label *lower_addr,
                 const orc_label *upper_addr, const orc_scalar *diag, const orc_scalar *upper,
                 const orc_scalar *lower, orc_label n_if, const orc_label *if_rows,
                 const orc_label *if_cols, const orc_scalar *if_bou, const orc_scalar *source,
                 orc_scalar *psi, orc_scalar tolerance, orc_scalar rel_tol, orc_label min_iter,
                 orc_label max_iter, orc_solve_result *result, orc_scalar *history,
                 orc_label history_cap);

/* Same for PBiCGStab::scalarSolve (asymmetric matrices).  history: the residual at every
 * convergence test (initial, then on s and on r in each iteration). */
int orc_foam_pbicgstab(orc_label n, orc_label n_faces, const orc_label *lower_addr,
                       const orc_label *upper_addr, const orc_scalar *diag, const orc_scalar *upper,
                       const orc_scalar *lower, orc_label n_if, const orc_label *if_rows,
                       const orc_label *if_cols, const orc_scalar *if_bou, const orc_scalar *source,
                       orc_scalar *psi, orc_scalar tolerance, orc_scalar rel_tol, orc_label min_iter,
                       orc_label max_iter, orc_solve_result *result, orc_scalar *history,
                       orc_label history_cap);

#ifdef __cplusplus
}
#endif
#endif
