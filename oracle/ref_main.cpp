// TEST INFRASTRUCTURE — headless driver for the UNMODIFIED reference solver.
// Compiled by oracle/Makefile against /root/reference/DFTAtom/{DFTAtom,PoissonSolver}.cpp where
// they lie (never copied into this repo).  Output: oracle/_ref/dftatom_ref (git-ignored).
// Replaces the wx worker-thread lambda of DFTAtomFrame.cpp:185-198 (the only caller of the solver).
//
// usage: dftatom_ref Z levels mixing rmax delta method(0=LDA,1=LSDA on the logarithmic grid - the product's only path,
//        DFTAtomFrame.cpp:190-195; 2=LDA, 3=LSDA on the uniform grid, DFTAtom.h:15,18: public but without a live caller, delta unused)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <limits>
#include <cmath>
#include "DFTAtom.h"

int main(int argc, char** argv)
{
    if (argc < 7) {
        std::fprintf(stderr, "usage: %s Z levels mixing rmax delta method\n", argv[0]);
        return 2;
    }
    const int Z = std::atoi(argv[1]);
    const int levels = std::atoi(argv[2]);
    const double mixing = std::atof(argv[3]);
    const double rmax = std::atof(argv[4]);
    const double delta = std::atof(argv[5]);
    const int method = std::atoi(argv[6]);
    if (method == 3) DFT::DFTAtom::CalculateUniformLSDA(Z, levels, mixing, rmax);
    else if (method == 2) DFT::DFTAtom::CalculateUniformLDA(Z, levels, mixing, rmax);
    else if (method) DFT::DFTAtom::CalculateNonUniformLSDA(Z, levels, mixing, rmax, delta);
    else DFT::DFTAtom::CalculateNonUniformLDA(Z, levels, mixing, rmax, delta);
    std::printf("\n");
    return 0;
}
