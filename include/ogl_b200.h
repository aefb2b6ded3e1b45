/* ogl_b200.h -- C ABI of libogl_b200.so, the B200 (sm_100a) sparse linear-solve
 * backend behind OGL's lduMatrix::solver plugin surface.
 *
 * This is the LOWER boundary of SURVEY.md section 8(b): what OGL's host C++
 * (HostMatrixWrapper, Persistent*, MatrixWrapper, Preconditioner,
 * StoppingCriterion, lduLduBase) calls INSTEAD of Ginkgo.  Each entry point
 * names the reference interface it replaces (paths relative to the OGL tree).
 *
 * Conventions
 *   - plain C: opaque handle, int status (0 = OGL_OK), no C++ types/exceptions
 *     cross the boundary; ogl_last_error() returns the message of the last
 *     failure on that handle (or of the last failed ogl_ctx_create when NULL).
 *   - label = int32_t, scalar = double (integration-tests.yml:13-14).
 *   - host pointers are caller-owned and only touched during the call; the
 *     library owns all device memory until ogl_ctx_destroy().
 *   - one context per (device, field) -- the analogue of the per-field registry
 *     objects "<field>_local_cols", "<field>_matrix", "<field>_rhs", ...
 *     (HostMatrix.H:21-65, CsrMatrixWrapper.H:251-260, lduLduBase.H:217-237).
 *     Calls on one handle are serialised by the caller (one thread per rank,
 *     as in the reference); work is stream-ordered internally.
 *   - there is NO CPU fallback: every compute entry point fails with
 *     OGL_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef OGL_B200_H
#define OGL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ogl_ctx ogl_ctx;

enum {
    OGL_OK = 0,
    OGL_ERR_INVALID = 1,   /* bad argument / call order                      */
    OGL_ERR_CUDA = 2,      /* CUDA runtime failure or no usable device       */
    OGL_ERR_NCCL = 3,      /* NCCL failure                                   */
    OGL_ERR_UNSUPPORTED = 4
};

#define OGL_NCCL_ID_BYTES 128

/* ---- executor / device ---------------------------------------------------
 * replaces ExecutorHandler (DevicePersistent/ExecutorHandler/ExecutorHandler.H:45-112,
 * :125-147: `executor cuda`, device = rank / ranksPerGPU % num_devices) and
 * DeviceIdGuard (DevicePersistent/DeviceIdGuard/DeviceIdGuard.H:15-42), plus the
 * two mpi::communicator objects (:140-144) which become one NCCL communicator.
 *
 * nccl_id: OGL_NCCL_ID_BYTES bytes produced by ogl_nccl_unique_id() on rank 0
 * and distributed by the caller (MPI_Bcast in OpenFOAM, torch.distributed
 * here); may be NULL when n_ranks == 1, and also when n_ranks > 1 if the caller
 * bootstraps the peer-memory windows itself (ogl_partition_export / _connect).
 * stream: a cudaStream_t to order all work on, or NULL for a private stream. */
int ogl_nccl_unique_id(void *out_id);
int ogl_device_count(int *count);
int ogl_ctx_create(int device_id, int rank, int n_ranks, const void *nccl_id,
                   void *stream, ogl_ctx **out);
int ogl_ctx_destroy(ogl_ctx *ctx);
const char *ogl_last_error(const ogl_ctx *ctx);
/* tuning knobs for experiments; unknown keys fail with OGL_ERR_INVALID, setting a key to its
 * current value keeps the cached iteration graph.  Kernel choice: "spmv_variant" (0 auto from the
 * row-length histogram, 1 stream, 2 thread/row, 3 warp/row, 4 TMA CSR, 5 warp tile, 6 pipelined
 * stream, 7 ELL, 8 merge-path = entry-balanced slices for very uneven row lengths), "ell_auto", "ell_coded" (pattern-coded ELL columns: 0 off, 1 auto, 2 force),
 * "ell_tma" (TMA-fed coded ELL: 0 off, 1 on, 2 auto by size), "tma_stages", "ell_minb", "ell_chunk",
 * "fuse_p" (CG p-update inside the ELL SpMV), "fused_pcg" (persistent CG loop kernel: 0/1/2 auto),
 * "gmres_persist", "tri_variant" (ILU / IC sweeps: 0 one launch per dependency level, 1 one
 * dependency-driven launch per sweep; read-only "tri_levels_lower" / "tri_levels_upper").  Loop control: "use_graph", "device_loop", "loop_iters", "chunk_iters",
 * "use_pdl".  Several ranks: "comm_mode" (0 auto, 1 NCCL, 2 peer memory), "fused_halo", "ghost_p".
 * Measurement: "profile_stride", "trace", "blas1_blocks", "stream_ctas", "l2_keep_mb". */
int ogl_set_option(ogl_ctx *ctx, const char *key, int64_t value);
int ogl_get_option(ogl_ctx *ctx, const char *key, int64_t *value);

/* ---- local sparsity pattern (device sort) ---------------------------------
 * replaces init_local_sparsity (HostMatrix/HostMatrixFreeFunctions.C:105-201)
 * and the cyclic-interface merge of init_local_sparsity_pattern
 * (HostMatrix/HostMatrix.C:468-589, collect_local_interface_indices :385-410).
 * lower_addr/upper_addr: lduAddr().lowerAddr()/upperAddr() (n_faces each).
 * iface_rows/iface_cols: the n_local_iface local (cyclic) couplings in
 * interface order (row = faceCell, col = neighbour patch faceCell).
 * Builds, on the device, row-major rows / cols / ldu_mapping and CSR row_ptrs,
 * bit-identical to the reference's host result; they stay resident and are
 * reused by every later solve (the reference re-uploads ldu_mapping each
 * solve, HostMatrix.C:699-701). */
int ogl_pattern_from_ldu(ogl_ctx *ctx, int32_t n_rows, int32_t n_faces,
                         int symmetric, const int32_t *lower_addr,
                         const int32_t *upper_addr, int32_t n_local_iface,
                         const int32_t *iface_rows, const int32_t *iface_cols);
int ogl_pattern_nnz(ogl_ctx *ctx, int64_t *local_nnz, int64_t *nonlocal_nnz);
/* download for inspection / parity (any pointer may be NULL);
 * PersistentArray::get_data on "<f>_local_rows/_cols/_ldu_map" */
int ogl_pattern_download(ogl_ctx *ctx, int32_t *rows, int32_t *cols,
                         int32_t *ldu_mapping, int32_t *row_ptrs);

/* ---- partition / communication pattern ------------------------------------
 * replaces PersistentPartition / PartitionInitFunctor::init
 * (DevicePersistent/Partition/Partition.H:57-70, :105-122:
 * localized_partition::build_from_blocked_recv + all_reduce of the local
 * size) fed by create_communication_pattern (HostMatrix.C:251-306).
 * target_ids ascending neighbour ranks, target_sizes, send_idxs concatenated
 * by target.  Collective over all ranks of the context's communicator. */
int ogl_partition_create(ogl_ctx *ctx, int32_t n_local, int32_t n_targets,
                         const int32_t *target_ids, const int32_t *target_sizes,
                         const int32_t *send_idxs);
int ogl_partition_sizes(ogl_ctx *ctx, int64_t *local_size, int64_t *global_size);
/* Host-driven bootstrap of the peer-memory windows, for contexts created with
 * n_ranks > 1 and nccl_id == NULL: the caller owns the communicator, as OGL's
 * host layer owns MPI_COMM_WORLD (DevicePersistent/ExecutorHandler/
 * ExecutorHandler.H:140-144, Partition.H:118-121 all_reduce of the local size).
 * After ogl_partition_create: export this rank's directory (an opaque blob;
 * blob == NULL queries its size), all-gather the blobs in rank order with the
 * host's communicator, connect, then run a host barrier before the first solve.
 * Works with several ranks per device (CUDA IPC), which NCCL does not. */
int ogl_partition_export(ogl_ctx *ctx, void *blob, int64_t capacity, int64_t *size);
int ogl_partition_connect(ogl_ctx *ctx, const void *blobs_all_ranks, int64_t n_blobs);

/* ---- non-local (halo) pattern ----------------------------------------------
 * replaces init_non_local_sparsity_pattern (HostMatrix.C:438-466) +
 * collect_cells_on_non_local_interface (:412-436).  face_cells: faceCells of
 * every processor interface concatenated in interface order; entry k couples
 * row face_cells[k] with recv-buffer slot k.  Sorted by row on the device with
 * ties in k order (stable; the reference's std::sort leaves ties unspecified). */
int ogl_nonlocal_pattern(ogl_ctx *ctx, int32_t n_halo, const int32_t *face_cells);
int ogl_nonlocal_pattern_download(ogl_ctx *ctx, int32_t *rows, int32_t *cols,
                                  int32_t *mapping);

/* ---- coefficient update (every solve) ---------------------------------------
 * replaces update_local_matrix_data (HostMatrix.C:592-705: H2D of
 * upper/lower/diag/local interface coeffs + row_gather through ldu_mapping),
 * update_non_local_matrix_data (:708-732), collect_interface_coeffs (:180-207,
 * the sign flip) and MatrixInitFunctor::update (CsrMatrixWrapper.H:74-136).
 * lower may be NULL for a symmetric matrix; local_iface_bou / nonlocal_bou are
 * the concatenated interfaceBouCoeffs (NOT negated; the library negates).
 * One staging upload + one gather kernel; no mapping re-upload, no D2D hop.
 * `scaling` multiplies every coefficient (documented intent of README.md:81). */
int ogl_values_update(ogl_ctx *ctx, const double *diag, const double *upper,
                      const double *lower, const double *local_iface_bou,
                      const double *nonlocal_bou, double scaling);
int ogl_values_download(ogl_ctx *ctx, double *local_vals, double *nonlocal_vals);

/* ---- vectors ------------------------------------------------------------------
 * replaces PersistentVector (DevicePersistent/Vector/Vector.H:52-83 init/update,
 * :144-167 copy_back).  b and x live on the DEVICE (the reference keeps them on
 * the host and clones per apply, lduLduBase.H:225,236).  x persists between
 * solves unless re-uploaded (updateInitGuess, lduLduBase.H:228-237). */
enum { OGL_VEC_B = 0, OGL_VEC_X = 1 };
int ogl_vector_upload(ogl_ctx *ctx, int which, const double *host, double scale);
int ogl_vector_download(ogl_ctx *ctx, int which, double *host);
int ogl_vector_fill(ogl_ctx *ctx, int which, double value);

/* ---- preconditioner -------------------------------------------------------------
 * replaces Preconditioner::init_preconditioner_impl("BJ"|"none") + wrap_schwarz
 * (Preconditioner/Preconditioner.H:47-64, :91-105, :342-344): Jacobi on the
 * LOCAL block only.  max_block_size 1 = inverse diagonal; 2..32 = block
 * detection + Gauss-Jordan inverses (Ginkgo jacobi::find_blocks/generate). */
/* ISAI: Ginkgo preconditioner::Isai<isai_type::spd> (Preconditioner/Preconditioner.H:225-242; needs an
 * SPD local matrix, i.e. `scaling -1` for OpenFOAM's pressure equation, README.md:101), GISAI:
 * Isai<isai_type::general> (:243-260); sparsityPower 1, rows up to 8 entries */
/* ILU: Ginkgo factorization::Ilu (exact ILU(0)) + preconditioner::Ilu<LowerTrs, UpperTrs>
 * (Preconditioner/Preconditioner.H:106-124); IC: factorization::Ic + preconditioner::Ic (:177-196,
 * needs an SPD local matrix like ISAI); IRILU: the ILU(0) factors with each triangular solve replaced
 * by 5 Richardson sweeps preconditioned with scalar Jacobi (:143-176).  All on the LOCAL block
 * (wrap_schwarz, :66-82).  The factorisation and the exact sweeps run level by level over the
 * dependency graph of the pattern, analysed once per mesh on the device.  ILUT / ICT (ParILUT / ParICT,
 * threshold-based and non-deterministic in the reference) are not built. */
enum { OGL_PRECOND_NONE = 0, OGL_PRECOND_BJ = 1, OGL_PRECOND_ISAI = 2, OGL_PRECOND_GISAI = 3,
       OGL_PRECOND_ILU = 4, OGL_PRECOND_IC = 5, OGL_PRECOND_IRILU = 6,
       /* Multigrid (Preconditioner/Preconditioner.H:261-341): Ginkgo multigrid::Pgm (deterministic)
        * aggregation, one V cycle per apply, Ir(2 sweeps, 0.9, scalar Jacobi) as pre- and
        * post-smoother, `coarseSolverIters` CG iterations on the coarsest level; options
        * "mg_max_levels" (maxLevels, 9), "mg_min_coarse_rows" (minCoarseRows, 10),
        * "mg_coarse_iters" (coarseSolverIters, 4), set before ogl_precond_setup; cycle v only */
       OGL_PRECOND_MULTIGRID = 7 };
int ogl_precond_setup(ogl_ctx *ctx, int kind, int32_t max_block_size,
                      int skip_sorting);
/* parity hook: the incomplete factors over the local CSR pattern ([nnz]; strictly lower part = L,
 * upper part incl. the diagonal = U; IC: lower part incl. the diagonal = L, upper part = L^T) */
int ogl_precond_factors_download(ogl_ctx *ctx, double *factors);
/* parity hooks for Multigrid: number of levels (the last one is the coarsest matrix), a level's
 * sizes (n_coarse = 0 on the coarsest level) and its CSR matrix / aggregate of every row */
int ogl_mg_levels(ogl_ctx *ctx, int32_t *n_levels);
int ogl_mg_level_info(ogl_ctx *ctx, int32_t level, int32_t *n, int32_t *nnz, int32_t *n_coarse);
int ogl_mg_level_download(ogl_ctx *ctx, int32_t level, int32_t *row_ptrs, int32_t *cols, double *vals,
                          int32_t *agg);
/* z = M^-1 r through the current preconditioner on host vectors of the local size (parity hook for
 * the triangular sweeps; any preconditioner kind) */
int ogl_precond_apply(ogl_ctx *ctx, const double *r_host, double *z_host);
/* parity hooks: block pointers and inverted blocks (row-major, concatenated) */
int ogl_precond_download(ogl_ctx *ctx, int32_t *n_blocks, int32_t *block_ptrs,
                         double *inv_blocks);

/* ---- solve -----------------------------------------------------------------------
 * replaces solver_gen->generate(A) + solver->apply(b, x) (lduLduBase.H:268-276)
 * i.e. gko::solver::{Cg,Bicgstab,Gmres}::apply with OGL's
 * OpenFOAMDistStoppingCriterion (StoppingCriterion/StoppingCriterion.C:71-151,
 * normFactor :32-69) evaluated on the device. */
enum { OGL_SOLVER_CG = 0, OGL_SOLVER_BICGSTAB = 1, OGL_SOLVER_GMRES = 2 };

typedef struct ogl_solve_params {
    int32_t solver;      /* OGL_SOLVER_*                                        */
    double tolerance;    /* StoppingCriterion.H:167                             */
    double rel_tol;      /* :168                                                */
    int32_t min_iter;    /* effective minIter (after adaptation, :199-209)      */
    int32_t max_iter;    /* caller doubles it for BiCGStab (:188)               */
    int32_t frequency;   /* effective evaluation frequency                      */
    int32_t krylov_dim;  /* GMRES restart length; <=0 -> 100 (Ginkgo default)   */
    int32_t export_res;  /* keep the residual history (`export`, :115-117)      */
} ogl_solve_params;

typedef struct ogl_solve_result {
    double init_residual;     /* StoppingCriterion::get_init_res_norm()          */
    double final_residual;    /* get_res_norm()                                  */
    double norm_factor;
    int32_t criterion_calls;  /* iter_ (get_num_iters())                         */
    int32_t n_iterations;     /* what solverPerformance reports (BiCGStab: /2)   */
    double solve_us;          /* device time of the iteration loop (CUDA events) */
    double resnorm_us;        /* time attributable to one criterion evaluation   */
    int64_t kernel_launches;  /* kernels of this library launched by the solve   */
    double spmv_us_avg;       /* sampled SpMV launch duration (profile_stride>0) */
    int32_t spmv_samples;
} ogl_solve_result;

int ogl_solve(ogl_ctx *ctx, const ogl_solve_params *params,
              ogl_solve_result *result);
int ogl_residual_history(ogl_ctx *ctx, double *out, int32_t capacity,
                         int32_t *written);

/* ---- building blocks exposed for parity tests and roofline measurement -----------
 * y = A x through the distributed operator (local CSR + halo exchange +
 * non-local block), gko::experimental::distributed::Matrix::apply. */
int ogl_spmv(ogl_ctx *ctx, const double *x_host, double *y_host);
/* `reps` back-to-back device SpMVs on resident vectors (optionally fused with
 * the <x,y> reduction CG needs next), timed with CUDA events on the context's
 * stream; *ms is the total. */
int ogl_spmv_bench(ogl_ctx *ctx, int32_t reps, int fused_dot, float *ms);
/* `iters` PCG iterations on the resident system without convergence exit
 * (roofline measurement of the whole iteration), timed with CUDA events. */
int ogl_pcg_bench(ogl_ctx *ctx, int32_t iters, float *ms);
int ogl_synchronize(ogl_ctx *ctx);
/* Diagnostics: with option `trace` 1 the iteration kernels log (tag << 48 | %globaltimer ns)
 * events -- launch start, last CTA arrived, local sums done, all-reduced, epilogue done; tags
 * 10.. p-update, 20.. SpMV, 30.. x/r-update -- into a device buffer.  Copies up to `cap` events
 * to the host and restarts the timeline.  No reference counterpart (nsys-style tooling). */
int ogl_trace_download(ogl_ctx *ctx, uint64_t *events, int64_t cap, int64_t *n_events);

/* HBM calibration kernels (roofline context for the numbers above), timed with
 * CUDA events; *gbs = bytes moved / time.  mode 0: copy (read + write, 128-bit),
 * 1: read-only sum, 2: read-only 8 B + 4 B streams (values + columns like CSR). */
int ogl_membench(ogl_ctx *ctx, int mode, int64_t n_doubles, int32_t reps, double *gbs);
/* Peer-memory primitives (multi-GPU, collective): mode 0 one all-reduce per
 * launch, 1 `reps` all-reduces inside one launch (device-side latency), 2 one
 * halo exchange (pack + consume) per repetition; *us = time per repetition. */
int ogl_commbench(ogl_ctx *ctx, int mode, int32_t reps, double *us);

/* ---- Matrix-Market export ---------------------------------------
 * replaces export_mtx / export_vec (common/common.C:31-58,
 * CsrMatrixWrapper.H:273-290, Vector.H:173-176): coordinate layout,
 * setprecision(15).  which: 0 local, 1 non-local, 2 rhs b, 3 the partition
 * side-car `<field>_partition.json` ({rank, n_ranks, n_local, target_ids,
 * target_sizes}: OGL's communication pattern, HostMatrix.C:251-306), which the
 * reference does not write and without which a decomposed dump cannot be read
 * back (ogl_b200/mtxio.py:import_decomposed). */
int ogl_export_mtx(ogl_ctx *ctx, int which, const char *path);

#ifdef __cplusplus
}
#endif
#endif /* OGL_B200_H */
