// fused pCN sweep instantiations for the rosenbrock likelihood (see pcn_sweep.cuh)
#include "pcn_sweep.cuh"
namespace asp {
int launch_sweep_rosenbrock(const SweepArgs& A, int g, int grid, cudaStream_t s) { return launch_sweep_like<ASP_DENSITY_ROSENBROCK>(A, g, grid, s); }
}  // namespace asp
