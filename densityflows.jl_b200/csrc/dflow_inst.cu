// One (HP, S[, REG]) instantiation of the chain kernels per object file: compile with
//   -DDFLOW_INST_FWD -DDFLOW_HP=16 -DDFLOW_S=2 -DDFLOW_REG=0      or      -DDFLOW_INST_GRAD2 -DDFLOW_HP=16 -DDFLOW_S=2
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

#ifdef DFLOW_CBANK  // this unit owns one constant bank (g_cbank)
}  // namespace dflow
#include <mutex>
namespace dflow {
// All users of the bank are totally ordered: (upload, kernel) pairs are enqueued under one lock, and every pair first
// waits on the event recorded after the previous pair, whatever stream that ran on.
static std::mutex g_cb_mu;
static cudaEvent_t g_cb_ev[64];

template <class Launch>
static cudaError_t with_bank(const void* chain, int chain_bytes, const float* staged, int stage_floats, cudaStream_t st,
                             Launch launch) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  const size_t w_off = (size_t)((chain_bytes + 15) / 16) * 16;
  if (w_off + (size_t)stage_floats * 4 > (size_t)CBANK_BYTES) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(g_cb_mu);
  if (!g_cb_ev[dev]) {
    e = cudaEventCreateWithFlags(&g_cb_ev[dev], cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  if ((e = cudaStreamWaitEvent(st, g_cb_ev[dev], 0)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(g_cbank, chain, chain_bytes, 0, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(g_cbank, staged, (size_t)stage_floats * 4, w_off, cudaMemcpyDeviceToDevice, st)) !=
      cudaSuccess)
    return e;
  if ((e = launch()) != cudaSuccess) return e;
  return cudaEventRecord(g_cb_ev[dev], st);
}
#endif

#ifdef DFLOW_INST_CFWD
template <>
cudaError_t launch_fwd_const_inst<DFLOW_HP, DFLOW_S>(const FwdArgs& a, unsigned grid, int nt, size_t smem, cudaStream_t st,
                                                     int stage_floats) {
  auto kern = chain_fwd_const_kernel<DFLOW_HP, DFLOW_S>;
  return with_bank(a.chain, a.chain_bytes, a.staged, stage_floats, st, [&]() -> cudaError_t {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<grid, nt, smem, st>>>(a);
    return cudaGetLastError();
  });
}
#endif

#ifdef DFLOW_INST_GRAD2
template <>
cudaError_t launch_grad2_inst<DFLOW_HP, DFLOW_S>(const GradArgs& a, unsigned grid, int nt, size_t smem, cudaStream_t st) {
  const bool vjp = a.zbar != nullptr || a.jbar != nullptr || a.xbar_out != nullptr || a.thbar_out != nullptr;
  auto kern = vjp ? chain_vjp2_kernel<DFLOW_HP, DFLOW_S> : chain_grad2_kernel<DFLOW_HP, DFLOW_S>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, nt, smem, st>>>(a);
  return cudaGetLastError();
}
#endif

}  // namespace dflow
