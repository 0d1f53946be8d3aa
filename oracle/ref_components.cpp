// TEST INFRASTRUCTURE — C-callable shims around the UNMODIFIED reference classes, so that the
// Python tests can call the reference's own component functions (SURVEY.md §8c "component-level
// oracle").  Built into oracle/_ref/libdftatom_ref.so by oracle/Makefile from the headers under
// /root/reference/DFTAtom (not copied).  Nothing in the product links against this.
#include <vector>
#include <algorithm>
#include <limits>
#include <cmath>
#include <math.h>
#include <cstring>
#include "Numerov.h"
#include "AufbauPrinciple.h"
#include "PoissonSolver.h"
#include "Integral.h"
#include "ExcCor.h"
#include "VWNExcCor.h"

using NumerovNU = DFT::Numerov<DFT::NumerovFunctionNonUniformGrid>;

extern "C" {

// Numerov.h:272-349
int ref_numerov_count_nodes(const double* V, int n_nodes, double delta, double rmax, int l, double E, int nodes_limit)
{
    DFT::Potential pot;
    pot.m_potentialValues.assign(V, V + n_nodes);
    NumerovNU num(pot, delta, rmax, n_nodes);
    int cnt = 0;
    num.SolveSchrodingerCountNodes(n_nodes - 1, l, E, n_nodes - 1, nodes_limit, cnt);
    return cnt;
}

// Numerov.h:204-270 (public, no caller in the reference)
int ref_numerov_count_from_nucleus(const double* V, int n_nodes, double delta, double rmax, int l, double E, int nodes_limit)
{
    DFT::Potential pot;
    pot.m_potentialValues.assign(V, V + n_nodes);
    NumerovNU num(pot, delta, rmax, n_nodes);
    int cnt = 0;
    num.SolveSchrodingerCountNodesFromNucleus(n_nodes - 1, l, E, n_nodes - 1, nodes_limit, cnt);
    return cnt;
}

// Numerov.h:351-401
double ref_numerov_solution_in_zero(const double* V, int n_nodes, double delta, double rmax, int l, double E)
{
    DFT::Potential pot;
    pot.m_potentialValues.assign(V, V + n_nodes);
    NumerovNU num(pot, delta, rmax, n_nodes);
    return num.SolveSchrodingerSolutionInZero(n_nodes - 1, l, E, n_nodes - 1);
}

// batched versions (one Numerov object, many (l,E) lanes) for speed
void ref_numerov_lanes(const double* V, int n_nodes, double delta, double rmax, int n_lanes,
                       const int* l, const double* E, const int* nodes_limit, double* y0, int* count)
{
    DFT::Potential pot;
    pot.m_potentialValues.assign(V, V + n_nodes);
    NumerovNU num(pot, delta, rmax, n_nodes);
    for (int k = 0; k < n_lanes; ++k) {
        if (y0) y0[k] = num.SolveSchrodingerSolutionInZero(n_nodes - 1, l[k], E[k], n_nodes - 1);
        if (count) {
            int cnt = 0;
            num.SolveSchrodingerCountNodes(n_nodes - 1, l[k], E[k], n_nodes - 1, nodes_limit[k], cnt);
            count[k] = cnt;
        }
    }
}

// Numerov.h:403-504
long ref_numerov_match(const double* V, int n_nodes, double delta, double rmax, int l, double E, double* psi)
{
    DFT::Potential pot;
    pot.m_potentialValues.assign(V, V + n_nodes);
    NumerovNU num(pot, delta, rmax, n_nodes);
    long int matchPoint = 0;
    std::vector<double> r = num.SolveSchrodingerMatchSolutionCompletely(n_nodes - 1, l, E, n_nodes - 1, matchPoint);
    std::memcpy(psi, r.data(), sizeof(double) * n_nodes);
    return matchPoint;
}

// The same three sweeps on the UNIFORM grid (Numerov<NumerovFunctionRegularGrid>, called as DFTAtom.cpp:234,242,266,299 do:
// startPoint = MaxR, steps = N - 1), one object, many (l, E) lanes
void ref_numerov_uniform_lanes(const double* V, int n_nodes, double rmax, int n_lanes, const int* l, const double* E,
                               const int* nodes_limit, double* y0, int* count)
{
    DFT::Potential pot;
    pot.m_potentialValues.assign(V, V + n_nodes);
    DFT::Numerov<DFT::NumerovFunctionRegularGrid> num(pot, 0, rmax, n_nodes);
    for (int k = 0; k < n_lanes; ++k) {
        if (y0) y0[k] = num.SolveSchrodingerSolutionInZero(rmax, l[k], E[k], n_nodes - 1);
        if (count) {
            int cnt = 0;
            num.SolveSchrodingerCountNodes(rmax, l[k], E[k], n_nodes - 1, nodes_limit[k], cnt);
            count[k] = cnt;
        }
    }
}

long ref_numerov_uniform_match(const double* V, int n_nodes, double rmax, int l, double E, double* psi)
{
    DFT::Potential pot;
    pot.m_potentialValues.assign(V, V + n_nodes);
    DFT::Numerov<DFT::NumerovFunctionRegularGrid> num(pot, 0, rmax, n_nodes);
    long int matchPoint = 0;
    std::vector<double> r = num.SolveSchrodingerMatchSolutionCompletely(rmax, l, E, n_nodes - 1, matchPoint);
    std::memcpy(psi, r.data(), sizeof(double) * n_nodes);
    return matchPoint;
}

// PoissonSolver.h:20-49
void ref_poisson_uniform(int levels, int Z, double rmax, const double* density, double* U)
{
    DFT::PoissonSolver ps(levels);
    const int n = DFT::PoissonSolver::GetNumberOfNodes(levels);
    std::vector<double> d(density, density + n);
    std::vector<double> u = ps.SolvePoissonUniform(Z, rmax, d);
    std::memcpy(U, u.data(), sizeof(double) * n);
}

// PoissonSolver.h:51-81
void ref_poisson_nonuniform(int levels, double delta, int Z, double rmax, const double* density, double* U)
{
    DFT::PoissonSolver ps(levels, delta);
    const int n = DFT::PoissonSolver::GetNumberOfNodes(levels);
    std::vector<double> d(density, density + n);
    std::vector<double> u = ps.SolvePoissonNonUniform(Z, rmax, d);
    std::memcpy(U, u.data(), sizeof(double) * n);
}

// VWNExcCor.h:73-128
void ref_vwn_lda(const double* rho, int n, double* vexc, double* eexcdif)
{
    std::vector<double> d(rho, rho + n);
    std::vector<double> v = DFT::VWNExchCor::Vexc(d);
    std::vector<double> e = DFT::VWNExchCor::eexcDif(d);
    std::memcpy(vexc, v.data(), sizeof(double) * n);
    std::memcpy(eexcdif, e.data(), sizeof(double) * n);
}

// ExcCor.h:27-95 (improved: 0 = ChachiyoExchCorParam, 1 = ChachiyoExchCorImprovedParam)
void ref_xc_chachiyo(const double* rho, int n, int improved, double* vexc, double* eexcdif)
{
    std::vector<double> d(rho, rho + n);
    std::vector<double> v = improved ? DFT::ChachiyoExchCor<DFT::ChachiyoExchCorImprovedParam>::Vexc(d) : DFT::ChachiyoExchCor<DFT::ChachiyoExchCorParam>::Vexc(d);
    std::vector<double> e = improved ? DFT::ChachiyoExchCor<DFT::ChachiyoExchCorImprovedParam>::eexcDif(d) : DFT::ChachiyoExchCor<DFT::ChachiyoExchCorParam>::eexcDif(d);
    std::memcpy(vexc, v.data(), sizeof(double) * n);
    std::memcpy(eexcdif, e.data(), sizeof(double) * n);
}

// VWNExcCor.h:134-312
void ref_vwn_lsda(const double* na, const double* nb, int n, double* va, double* vb, double* vexc, double* eexcdif)
{
    std::vector<double> a(na, na + n), b(nb, nb + n), xa, xb;
    std::vector<double> v = DFT::VWNExchCor::Vexc(a, b, xa, xb);
    std::vector<double> e = DFT::VWNExchCor::eexcDif(a, b);
    std::memcpy(va, xa.data(), sizeof(double) * n);
    std::memcpy(vb, xb.data(), sizeof(double) * n);
    std::memcpy(vexc, v.data(), sizeof(double) * n);
    std::memcpy(eexcdif, e.data(), sizeof(double) * n);
}

// Integral.h:50-73
double ref_simpson38(double delta, const double* v, int n)
{
    std::vector<double> d(v, v + n);
    return DFT::Integral::Simpson38(delta, d);
}

// Integral.h:11-155: 0 Trapezoid, 1 SimpsonOneThird, 2 Simpson38, 3 Boole, 4 Romberg
double ref_integrate(int rule, double delta, const double* v, int n)
{
    std::vector<double> d(v, v + n);
    switch (rule) {
        case 0: return DFT::Integral::Trapezoid(delta, d);
        case 1: return DFT::Integral::SimpsonOneThird(delta, d);
        case 2: return DFT::Integral::Simpson38(delta, d);
        case 3: return DFT::Integral::Boole(delta, d);
        default: return DFT::Integral::Romberg(delta, d);
    }
}

// AufbauPrinciple.h:36-75 + the driver's sort (DFTAtom.cpp:367); out = triples (N0, L, occ); returns count
int ref_aufbau(int Z, int* out, int max_levels)
{
    std::vector<DFT::Subshell> lv = DFT::AufbauPrinciple::GetSubshells(Z);
    std::sort(lv.begin(), lv.end());
    int k = 0;
    for (const auto& s : lv) {
        if (k >= max_levels) break;
        out[3 * k] = s.m_N; out[3 * k + 1] = s.m_L; out[3 * k + 2] = s.m_nrElectrons;
        ++k;
    }
    return k;
}

}
