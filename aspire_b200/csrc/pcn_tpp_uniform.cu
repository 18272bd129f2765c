// particle-per-thread pCN sweep instantiations for the uniform likelihood (see pcn_sweep_tpp.cuh)
#include "pcn_sweep_tpp.cuh"
namespace asp {
int launch_tpp_uniform(const SweepArgs& A, cudaStream_t s) { return launch_tpp_like<ASP_DENSITY_UNIFORM>(A, s); }
}  // namespace asp
