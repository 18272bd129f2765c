// particle-per-thread tpCN sweep instantiations for the gaussian likelihood (see pcn_sweep_tpp.cuh)
#include "pcn_sweep_tpp.cuh"
namespace asp {
int launch_tpp_t_gaussian(const SweepArgs& A, cudaStream_t s) { return launch_tpp_like<ASP_DENSITY_GAUSSIAN, true>(A, s); }
}  // namespace asp
