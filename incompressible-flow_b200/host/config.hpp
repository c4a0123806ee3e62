// config.hpp -- compile-time constants of the hot path.  The reference generates this file with CMake
// [REF config.hpp.in]; only the two viscosities are on the IMPES path [REF config.hpp.in:37,43].
#pragma once
#if defined(IMPES_WITH_MFEM)
#include "mfem.hpp"
#else
#include "mfem_compat.hpp"
#include "options_parser.hpp"
#endif

const double MU_W = 1.0;  // water viscosity
const double MU_O = 1.0;  // oil viscosity
