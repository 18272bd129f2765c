// particle-per-thread pCN sweep instantiations for the gmm likelihood (see pcn_sweep_tpp.cuh)
#include "pcn_sweep_tpp.cuh"
namespace asp {
int launch_tpp_gmm(const SweepArgs& A, cudaStream_t s) { return launch_tpp_like<ASP_DENSITY_GMM>(A, s); }
}  // namespace asp
