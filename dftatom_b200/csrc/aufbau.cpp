// Host-side occupation rules (integer logic, runs once per atom).
// Replaces AufbauPrinciple::GetSubshells (reference DFTAtom/AufbauPrinciple.h:36-75, exceptions :101-117,
// :129-142) + the driver's sort by (n,l) (DFTAtom.cpp:367) and DFTAtom::InitializeLevels (DFTAtom.cpp:611-638).
#include <algorithm>
#include <vector>
#include "internal.h"

namespace dft {

namespace {

struct Shell { int n0, l, occ; };

// f-block / Lr adjustments.  The reference applies the same adjustment twice around the "electrons left"
// clamp (AufbauPrinciple.h:53,59); the transition-metal rules (:78-99) exist there but are never invoked,
// so Cr comes out 3d4 4s2 — reproduced on purpose.
int adjust(int occ, int Z, int n0, int l)
{
    if (l == 3) {
        const bool la_ce_gd = (Z == 57 || Z == 58 || Z == 64);
        if (la_ce_gd && n0 == 3) return occ - 1;
        if (n0 == 4) {
            if (Z == 89 || Z == 90) return 0;
            if (Z == 91 || Z == 92 || Z == 93 || Z == 96) return occ - 1;
        }
    } else if (Z == 103 && n0 == 5 && l == 2) {
        return 0;
    }
    return occ;
}

std::vector<Shell> fill(int Z)
{
    std::vector<Shell> v;
    int placed = 0;
    for (int sum = 0; sum < 10 && placed != Z; ++sum)
        for (int n0 = 0; n0 <= sum && placed != Z; ++n0) {
            const int l = sum - n0;
            if (l > n0) continue;
            int occ = adjust(2 * (2 * l + 1), Z, n0, l);
            occ = std::min(occ, Z - placed);
            occ = adjust(occ, Z, n0, l);
            if (occ > 0) { v.push_back({ n0, l, occ }); placed += occ; }
        }
    std::sort(v.begin(), v.end(), [](const Shell& a, const Shell& b) { return a.n0 != b.n0 ? a.n0 < b.n0 : a.l < b.l; });
    return v;
}

dftatom_level to_level(const Shell& s)
{
    dftatom_level L;
    L.n = s.n0 + 1; L.l = s.l; L.occ = s.occ; L.nodes = s.n0 - s.l; L.E = 0.;
    return L;
}

}  // namespace

int aufbau(int Z, dftatom_level* out, int max_out)
{
    if (Z < 1 || Z > 118) return DFTATOM_E_BAD_OPTION;
    const std::vector<Shell> v = fill(Z);
    if ((int)v.size() > max_out) return DFTATOM_E_ARG;
    for (size_t k = 0; k < v.size(); ++k) out[k] = to_level(v[k]);
    return (int)v.size();
}

int split_spin(int Z, dftatom_level* a, int* na, dftatom_level* b, int* nb, int* ea, int* eb)
{
    dftatom_level all[DFTATOM_MAX_LEVELS];
    const int n = aufbau(Z, all, DFTATOM_MAX_LEVELS);
    if (n < 0) return n;
    int ka = 0, kb = 0, n_alpha = 0;
    for (int k = 0; k < n; ++k) {
        const int cap = 2 * all[k].l + 1;                 // getMaxNrAlphaElectrons, AufbauPrinciple.h:26-29
        const int up = std::min(all[k].occ, cap);
        a[ka] = all[k]; a[ka].occ = up; ++ka;
        n_alpha += up;
        if (all[k].occ - up > 0) { b[kb] = all[k]; b[kb].occ = all[k].occ - up; ++kb; }
    }
    *na = ka; *nb = kb; *ea = n_alpha; *eb = Z - n_alpha;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Sharding of a batch of independent atoms over G GPUs (one process per GPU, no collective).
// Cost of an atom = (spin) orbitals x expected SCF steps.  Which atoms take 60-150 steps instead of ~35 is physics, not Z: the
// NODELESS compact d / f shells (3d, 4f: n = l + 1) slosh under linear mixing when they are nearly full (Cu, Zn: 89 / 69 steps,
// Ho .. Yb: 77 .. 100 steps of the reference's sweep; their 4d / 5d / 5f counterparts do not).  Only used to balance
// shards - a wrong estimate costs time, never results.
// ---------------------------------------------------------------------------------------------------------
double estimate_cost(int Z, int method)
{
    if (Z < 1 || Z > 118) return 0.;
    const std::vector<Shell> v = fill(Z);
    int n_orb = 0;
    for (const Shell& s : v) n_orb += (method && s.occ > 2 * s.l + 1) ? 2 : 1;
    // the subshells in Madelung (filling) order: the one filled last, and the one before it
    std::vector<const Shell*> order;
    for (const Shell& s : v) order.push_back(&s);
    std::sort(order.begin(), order.end(), [](const Shell* a, const Shell* b) { return a->n0 + a->l != b->n0 + b->l ? a->n0 + a->l < b->n0 + b->l : a->n0 < b->n0; });
    double steps = method ? 50. : 36.;
    for (size_t q = order.size() >= 2 ? order.size() - 2 : 0; q < order.size(); ++q) {
        const Shell& s = *order[q];
        if (s.l < 1 || s.n0 != s.l) continue;                   // nodeless: n = l + 1 (n0 is 0-based)
        const bool is_last = q + 1 == order.size();
        const double x = (double)s.occ / (double)(2 * (2 * s.l + 1));
        if (x > 0.7) steps += (method ? 100. : 65.) * (x - 0.7) / 0.3 * (s.l >= 2 ? 1. : 0.2) * (is_last ? 1. : 0.4);
    }
    return (double)n_orb * steps;
}

// longest-processing-time-first: atoms in order of decreasing cost, each to the least loaded rank; deterministic
int partition_atoms(const int* Z, const int* method, int n, int n_ranks, int* rank_of)
{
    if (!Z || !rank_of || n < 0 || n_ranks < 1) return DFTATOM_E_ARG;
    std::vector<double> cost(n);
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) { cost[i] = estimate_cost(Z[i], method ? method[i] : 0); order[i] = i; }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<double> load(n_ranks, 0.);
    for (int i : order) {
        int r = 0;
        for (int q = 1; q < n_ranks; ++q) if (load[q] < load[r]) r = q;
        rank_of[i] = r;
        load[r] += cost[i];
    }
    return 0;
}

}  // namespace dft
