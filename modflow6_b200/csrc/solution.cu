// solution.cu -- placeholder, replaced by the device-resident formulate path
#include "solver.cuh"
using namespace mf6;
extern "C" {
#define NOTYET(name) return guard([&] { MF6_REQUIRE(false, name ": not implemented yet"); })
int mf6gpu_solution_create(const mf6gpu_gwf_model *, const mf6gpu_sln_settings *, const mf6gpu_ims_settings *, mf6gpu_solution **) { NOTYET("solution_create"); }
int mf6gpu_solution_destroy(mf6gpu_solution *) { return 0; }
int mf6gpu_solution_set_packages(mf6gpu_solution *, int32_t, const mf6gpu_bnd_package *) { NOTYET("solution_set_packages"); }
int mf6gpu_solution_timestep(mf6gpu_solution *, int32_t, int32_t, double, int32_t, mf6gpu_step_report *) { NOTYET("solution_timestep"); }
int mf6gpu_solution_formulate(mf6gpu_solution *, int32_t, double, int32_t) { NOTYET("solution_formulate"); }
int mf6gpu_solution_get_x(mf6gpu_solution *, double *) { NOTYET("solution_get_x"); }
int mf6gpu_solution_set_x(mf6gpu_solution *, const double *) { NOTYET("solution_set_x"); }
int mf6gpu_solution_get_amat(mf6gpu_solution *, double *) { NOTYET("solution_get_amat"); }
int mf6gpu_solution_get_rhs(mf6gpu_solution *, double *) { NOTYET("solution_get_rhs"); }
int mf6gpu_solution_get_flowja(mf6gpu_solution *, double *) { NOTYET("solution_get_flowja"); }
int mf6gpu_solution_get_condsat(mf6gpu_solution *, double *) { NOTYET("solution_get_condsat"); }
mf6gpu_solver *mf6gpu_solution_solver(mf6gpu_solution *) { return nullptr; }
}
