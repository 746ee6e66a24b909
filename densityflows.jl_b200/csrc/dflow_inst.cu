// One (HP, S[, REG]) instantiation of the chain kernels per object file: compile with
//   -DDFLOW_INST_FWD -DDFLOW_HP=16 -DDFLOW_S=2 -DDFLOW_REG=0|1      or      -DDFLOW_INST_GRAD -DDFLOW_HP=16
#include "dflow_chain_kernels.cuh"
#include "dflow_grad_kernel.cuh"

namespace dflow {

#ifdef DFLOW_INST_FWD
template <>
cudaError_t launch_fwd_inst<DFLOW_HP, DFLOW_S, (DFLOW_REG != 0)>(const FwdArgs& a, unsigned grid, int nt, size_t smem,
                                                                cudaStream_t st) {
  auto kern = chain_fwd_kernel<DFLOW_HP, DFLOW_S, (DFLOW_REG != 0)>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, nt, smem, st>>>(a);
  return cudaGetLastError();
}
#endif

#ifdef DFLOW_INST_GRAD
template <>
cudaError_t launch_grad_inst<DFLOW_HP>(const GradArgs& a, unsigned grid, int nt, size_t smem, cudaStream_t st) {
  auto kern = chain_grad_kernel<DFLOW_HP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, nt, smem, st>>>(a);
  return cudaGetLastError();
}
#endif

#ifdef DFLOW_INST_GRAD2
template <>
cudaError_t launch_grad2_inst<DFLOW_HP, DFLOW_S>(const GradArgs& a, unsigned grid, int nt, size_t smem, cudaStream_t st) {
  auto kern = chain_grad2_kernel<DFLOW_HP, DFLOW_S>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, nt, smem, st>>>(a);
  return cudaGetLastError();
}
#endif

}  // namespace dflow
