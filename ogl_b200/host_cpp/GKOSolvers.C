// Run-time selection table entries, as in Solver/CG/GKOCG.C:14-17 (symmetric
// only), Solver/BiCGStab/GKOBiCGStab.C:14-20 and Solver/GMRES/GKOGMRES.C:14-20
// (symmetric + asymmetric); plus the out-of-line pieces of the shim.
#include "GKOSolvers.H"

namespace Foam {

const dictionary dictionary::null{};

void dictionary::parse(const std::string &text)
{
    // `key value;` and `key { ... }`, whitespace separated; enough for fvSolution solver dicts
    size_t i = 0;
    const size_t n = text.size();
    auto skip_ws = [&]() {
        while (i < n && std::isspace(static_cast<unsigned char>(text[i]))) ++i;
    };
    while (true) {
        skip_ws();
        if (i >= n) break;
        size_t k0 = i;
        while (i < n && !std::isspace(static_cast<unsigned char>(text[i])) && text[i] != '{' && text[i] != ';') ++i;
        const word key = text.substr(k0, i - k0);
        skip_ws();
        if (i < n && text[i] == '{') {
            int depth = 1;
            size_t b0 = ++i;
            while (i < n && depth > 0) {
                if (text[i] == '{') ++depth;
                if (text[i] == '}') --depth;
                ++i;
            }
            subs_[key].parse(text.substr(b0, i - b0 - 1));
        } else {
            size_t v0 = i;
            while (i < n && text[i] != ';') ++i;
            word val = text.substr(v0, i - v0);
            while (!val.empty() && std::isspace(static_cast<unsigned char>(val.back()))) val.pop_back();
            if (!key.empty()) entries_[key] = val;
            if (i < n) ++i;
        }
    }
}

lduMatrix::solver::addsymMatrixConstructorToTable<GKOCG> addGKOCGSymMatrixConstructorToTable_;

lduMatrix::solver::addsymMatrixConstructorToTable<GKOBiCGStab> addGKOBiCGStabSymMatrixConstructorToTable_;
lduMatrix::solver::addasymMatrixConstructorToTable<GKOBiCGStab> addGKOBiCGStabAsymMatrixConstructorToTable_;

lduMatrix::solver::addsymMatrixConstructorToTable<GKOGMRES> addGKOGMRESSymMatrixConstructorToTable_;
lduMatrix::solver::addasymMatrixConstructorToTable<GKOGMRES> addGKOGMRESAsymMatrixConstructorToTable_;

}  // namespace Foam
