// Test kernels for the branch-free IEEE-correct division / square root of the bit-exact build (om_runtime.cuh:
// om_rcp_rn_seq / om_div_rn / om_sqrt_rn) against the compiler's div.rn.f64 / sqrt.rn.f64 on the same operands.
// Built by __graft_entry__.build() into tests/cuda/_build/libdivsqrt_check.so; driven by tests/test_gpu_divsqrt.py.
#include "om_runtime.cuh"

__global__ void div_kernel(const double* a, const double* b, double* ours, double* ieee, unsigned* bad, long long n) {
  // bad[0]: some result was NaN / Inf (om_state_bad); bad[1]: number of elements whose `slow` flag came up
  unsigned flag = 0u, nslow = 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    bool slow = false;
    const double y = om_rcp_rn_seq(b[i], slow);
    ours[i] = om_div_rn(a[i], b[i], y, slow);          // always the branch-free sequence (a generated scope would redo slow cells)
    flag |= om_state_bad(ours[i]);
    nslow += slow ? 1u : 0u;
    ieee[i] = __ddiv_rn(a[i], b[i]);
  }
  if (flag) atomicOr(bad, 1u);
  if (nslow) atomicAdd(bad + 1, nslow);
}
__global__ void sqrt_kernel(const double* x, double* ours, double* ieee, unsigned* bad, long long n) {
  unsigned flag = 0u, nslow = 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    bool slow = false;
    ours[i] = om_sqrt_rn(x[i], slow);
    flag |= om_state_bad(ours[i]);
    nslow += slow ? 1u : 0u;
    ieee[i] = __dsqrt_rn(x[i]);
  }
  if (flag) atomicOr(bad, 1u);
  if (nslow) atomicAdd(bad + 1, nslow);
}
// the fast_math build's division / square root (om_frcp + om_fdiv_r, om_fsqrt): not IEEE, checked for their error bound
__global__ void fast_kernel(const double* a, const double* b, double* quot, double* root, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    quot[i] = om_fdiv_r(a[i], b[i], om_frcp(b[i]));
    root[i] = om_fsqrt(a[i]);
  }
}
extern "C" int om_check_fast(const void* a, const void* b, void* quot, void* root, long long n, void* stream) {
  fast_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((const double*)a, (const double*)b, (double*)quot, (double*)root, n);
  return (int)cudaGetLastError();
}
extern "C" int om_check_div(const void* a, const void* b, void* ours, void* ieee, void* bad, long long n, void* stream) {
  div_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((const double*)a, (const double*)b, (double*)ours, (double*)ieee, (unsigned*)bad, n);
  return (int)cudaGetLastError();
}
extern "C" int om_check_sqrt(const void* x, void* ours, void* ieee, void* bad, long long n, void* stream) {
  sqrt_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((const double*)x, (double*)ours, (double*)ieee, (unsigned*)bad, n);
  return (int)cudaGetLastError();
}
