#!/usr/bin/env bash
# Build libaspire_b200.so (sm_100a only) in-tree: aspire_b200/lib/libaspire_b200.so
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${ASP_LIB_OUT:-${HERE}/../lib}"
mkdir -p "${OUT}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr)
SRCS=(core.cu reweight.cu scan_exact.cu search_gather.cu sorted_resample.cu pcn.cu pcn_sweep_gaussian.cu pcn_sweep_rosenbrock.cu pcn_sweep_gmm.cu pcn_sweep_uniform.cu pcn_tpp_gaussian.cu pcn_tpp_rosenbrock.cu pcn_tpp_gmm.cu pcn_tpp_uniform.cu pcn_tpp_t_gaussian.cu pcn_tpp_t_rosenbrock.cu pcn_tpp_t_gmm.cu pcn_tpp_t_uniform.cu pcn_eval.cu student_t.cu rng.cu select.cu flow.cu flow_move.cu flow_fit.cu transform.cu)
OBJS=()
pids=()
for s in "${SRCS[@]}"; do
  o="${OUT}/${s%.cu}.o"
  OBJS+=("${o}")
  if [[ ! -f "${o}" || "${HERE}/${s}" -nt "${o}" || -n "$(find "${HERE}" "${HERE}/../../include" \( -name '*.cuh' -o -name '*.h' \) -newer "${o}" 2>/dev/null | head -1)" ]]; then
    "${NVCC}" "${FLAGS[@]}" ${ASP_NVCC_EXTRA:-} -c "${HERE}/${s}" -o "${o}" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "${p}" ]] && wait "${p}"; done
"${NVCC}" -shared -o "${OUT}/libaspire_b200.so" "${OBJS[@]}" -gencode arch=compute_100a,code=sm_100a
echo "built ${OUT}/libaspire_b200.so"
