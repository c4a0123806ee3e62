// solver_context.hpp -- the state the reference re-creates on every call (assembled matrices,
// preconditioners) lives, in the GPU build, in one impes_handle per Param.  The reference's solver
// functions are stateless [REF src/pressure_solver.cpp:119-124]; the adapter keeps that contract by
// looking the handle up from the Param address (created on first use, destroyed at exit or by
// impes_host_release).
#pragma once
#include "../../include/impes_b200.h"

struct Param;

impes_handle impes_host_context(const Param& param);
void impes_host_release(const Param& param);
