// C entry points that let the test-suite drive the C++ plugin layer exactly the
// way OpenFOAM would: build an lduMatrix + interfaces from flat arrays, select
// the solver from the run-time table by the fvSolution dictionary, call solve().
#include <cstring>

#include "GKOSolvers.H"

using namespace Foam;

namespace {
struct Case {
    lduMesh mesh;
    std::unique_ptr<lduAddressing> addr;
    std::unique_ptr<lduMatrix> matrix;
    std::vector<std::unique_ptr<lduInterface>> ifaces;
    std::vector<std::unique_ptr<lduInterfaceField>> fields;
    lduInterfaceFieldPtrsList ptrs;
    FieldField<Field, scalar> bou, intc;
};
thread_local std::string g_err;
}  // namespace

extern "C" {

const char *foamshim_last_error() { return g_err.c_str(); }

void foamshim_set_parallel(int rank, int n_ranks, const unsigned char *nccl_id)
{
    Pstream::parRunRef() = n_ranks > 1;
    Pstream::rankRef() = rank;
    Pstream::nProcsRef() = n_ranks;
    if (nccl_id) Pstream::ncclId().assign(nccl_id, nccl_id + OGL_NCCL_ID_BYTES);
}

// kinds: 0 processor, 1 cyclic, 2 cyclicAMI; nbr: neighbour rank (processor) or
// neighbour patch index (cyclic)
void *foamshim_case_create(int n, int n_faces, const int *lower, const int *upper, int n_ifaces,
                           const int *kinds, const int *nbr, const int *sizes, const int *face_cells)
{
    auto *c = new Case();
    c->addr.reset(new lduAddressing(labelList(lower, lower + n_faces), labelList(upper, upper + n_faces)));
    size_t off = 0;
    std::vector<const lduInterface *> patches;
    for (int i = 0; i < n_ifaces; ++i) {
        labelList fc(face_cells + off, face_cells + off + sizes[i]);
        off += sizes[i];
        if (kinds[i] == 0) c->ifaces.emplace_back(new processorLduInterface(fc, nbr[i]));
        else if (kinds[i] == 1) c->ifaces.emplace_back(new cyclicFvPatch(fc, nbr[i]));
        else c->ifaces.emplace_back(new cyclicAMIFvPatch(fc));
        c->fields.emplace_back(new lduInterfaceField(*c->ifaces.back()));
        c->ptrs.push_back(c->fields.back().get());
        patches.push_back(c->ifaces.back().get());
    }
    c->addr->setPatches(patches);
    c->bou.resize(n_ifaces);
    c->intc.resize(n_ifaces);
    (void)n;
    return c;
}

void foamshim_case_destroy(void *h) { delete static_cast<Case *>(h); }

int foamshim_registry_size(void *h) { return static_cast<Case *>(h)->mesh.thisDb().size(); }

// (re)sets the coefficients -- a new "time step" on the same mesh
void foamshim_case_set_coeffs(void *h, int n, int n_faces, const double *diag, const double *upper,
                              const double *lower_or_null, const double *bou_concat)
{
    auto *c = static_cast<Case *>(h);
    c->matrix.reset(new lduMatrix(c->mesh, *c->addr, scalarField(diag, diag + n),
                                  scalarField(upper, upper + n_faces),
                                  lower_or_null ? scalarField(lower_or_null, lower_or_null + n_faces)
                                                : scalarField()));
    size_t off = 0;
    for (size_t i = 0; i < c->ifaces.size(); ++i) {
        const size_t sz = static_cast<size_t>(c->ifaces[i]->faceCells().size());
        c->bou[i] = scalarField(bou_concat + off, bou_concat + off + sz);
        c->intc[i] = scalarField(sz, 0.0);
        off += sz;
    }
}

// lduMatrix::solver::New(fieldName, ...)->solve(psi, source); returns 0 or 1 (FatalError)
int foamshim_solve(void *h, const char *field_name, const char *dict_text, int n, double *psi,
                   const double *source, char *solver_name_out, int name_cap, double *init_res,
                   double *final_res, int *n_iter)
{
    auto *c = static_cast<Case *>(h);
    try {
        dictionary controls{std::string(dict_text)};
        scalarField psi_f(psi, psi + n), src_f(source, source + n);
        auto solver = lduMatrix::solver::New(field_name, *c->matrix, c->bou, c->intc, c->ptrs, controls);
        solverPerformance perf = solver->solve(psi_f, src_f);
        std::memcpy(psi, psi_f.data(), sizeof(double) * n);
        if (solver_name_out && name_cap > 0) {
            std::strncpy(solver_name_out, perf.solverName().c_str(), name_cap - 1);
            solver_name_out[name_cap - 1] = 0;
        }
        *init_res = perf.initialResidual();
        *final_res = perf.finalResidual();
        *n_iter = perf.nIterations();
        return 0;
    } catch (const FatalErrorException &e) {
        g_err = e.what();
        return 1;
    }
}

}  // extern "C"
