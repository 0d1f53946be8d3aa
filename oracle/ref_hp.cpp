// TEST INFRASTRUCTURE — full-precision (17 significant digits) build of the UNMODIFIED reference
// SCF driver: the standard headers are included first so that the macro below only rewrites the
// reference's own `std::setprecision(6)` calls (DFTAtom.cpp:472,556).  The reference file is
// included from its mounted path, never copied.
#include <iostream>
#include <iomanip>
#include <vector>
#include <algorithm>
#include <fstream>
#include <sstream>
#include <cassert>
#include <limits>
#include <cmath>
#include <math.h>
#define setprecision(x) setprecision(17)
#include "DFTAtom.cpp"
