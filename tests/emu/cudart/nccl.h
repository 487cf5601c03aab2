// TEST INFRASTRUCTURE ONLY — single-process stand-in for the NCCL calls of the generated host class (ncclCommInitAll,
// grouped ncclSend / ncclRecv, ncclAllReduce of one scalar slot) over the host-memory "devices" of
// tests/emu/cudart/cuda_runtime_api.h.  Operations are queued between ncclGroupStart and ncclGroupEnd and executed at the
// end of the group: every recv is paired with the oldest unmatched send of its peer, all-reduces are folded over the ranks
// in rank order.  Lets the slab decomposition of the C++ class (ghost-row geometry, deferred all-reduce) run without GPUs.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "cuda_runtime_api.h"

typedef int ncclResult_t;
enum { ncclSuccess = 0, ncclInvalidUsage = 5 };
typedef enum { ncclChar = 0, ncclInt32 = 2, ncclInt64 = 4, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
struct ncclCommEmu { int rank, size; };
typedef ncclCommEmu* ncclComm_t;

namespace om_nccl_emu {
struct P2P { bool send; char* buf; size_t bytes; int me, peer; bool done; };
struct AR { const void* in; void* out; size_t count; ncclDataType_t t; ncclRedOp_t op; int rank, size; };
inline std::vector<P2P>& p2p() { static std::vector<P2P> v; return v; }
inline std::vector<AR>& ars() { static std::vector<AR> v; return v; }
inline int& depth() { static int d = 0; return d; }
inline size_t elem(ncclDataType_t t) { return t == ncclChar ? 1 : (t == ncclInt32 || t == ncclFloat32) ? 4 : 8; }
template <class T> inline void fold(std::vector<AR*>& g) {
  for (size_t i = 0; i < g[0]->count; ++i) {
    T acc = ((const T*)g[0]->in)[i];
    for (size_t r = 1; r < g.size(); ++r) {
      const T v = ((const T*)g[r]->in)[i];
      acc = g[0]->op == ncclSum ? (T)(acc + v) : g[0]->op == ncclMin ? (v < acc ? v : acc) : (acc < v ? v : acc);
    }
    for (auto* a : g) ((T*)a->out)[i] = acc;
  }
}
inline ncclResult_t flush() {
  auto& q = p2p();
  for (auto& r : q) if (!r.send) {
    bool ok = false;
    for (auto& s : q) if (s.send && !s.done && s.me == r.peer && s.peer == r.me) {
      if (s.bytes != r.bytes) return ncclInvalidUsage;
      std::memcpy(r.buf, s.buf, r.bytes); s.done = true; ok = true; break;
    }
    if (!ok) return ncclInvalidUsage;
  }
  for (auto& s : q) if (s.send && !s.done) return ncclInvalidUsage;       // a send nobody receives would hang real NCCL
  q.clear();
  auto& a = ars();
  while (!a.empty()) {                 // the k-th all-reduce of every rank belongs together
    const int size = a[0].size;
    std::vector<AR*> g(size, nullptr);
    std::vector<size_t> used;
    for (size_t i = 0; i < a.size(); ++i) if (!g[a[i].rank]) { g[a[i].rank] = &a[i]; used.push_back(i); }
    for (auto* x : g) if (!x) return ncclInvalidUsage;
    switch (g[0]->t) {
      case ncclInt32: fold<int32_t>(g); break;
      case ncclInt64: fold<int64_t>(g); break;
      case ncclFloat32: fold<float>(g); break;
      case ncclFloat64: fold<double>(g); break;
      default: return ncclInvalidUsage;
    }
    for (size_t k = used.size(); k-- > 0;) a.erase(a.begin() + used[k]);
  }
  return ncclSuccess;
}
}

static inline const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL usage error"; }
static inline ncclResult_t ncclCommInitAll(ncclComm_t* comms, int n, const int*) {
  for (int i = 0; i < n; ++i) comms[i] = new ncclCommEmu{i, n};
  return ncclSuccess;
}
static inline ncclResult_t ncclCommDestroy(ncclComm_t c) { delete c; return ncclSuccess; }
static inline ncclResult_t ncclGroupStart() { ++om_nccl_emu::depth(); return ncclSuccess; }
static inline ncclResult_t ncclGroupEnd() { return --om_nccl_emu::depth() == 0 ? om_nccl_emu::flush() : ncclSuccess; }
static inline ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  om_nccl_emu::p2p().push_back({true, (char*)buf, count * om_nccl_emu::elem(t), c->rank, peer, false});
  return om_nccl_emu::depth() ? ncclSuccess : om_nccl_emu::flush();
}
static inline ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  om_nccl_emu::p2p().push_back({false, (char*)buf, count * om_nccl_emu::elem(t), c->rank, peer, false});
  return om_nccl_emu::depth() ? ncclSuccess : om_nccl_emu::flush();
}
static inline ncclResult_t ncclAllReduce(const void* in, void* out, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c, cudaStream_t) {
  om_nccl_emu::ars().push_back({in, out, count, t, op, c->rank, c->size});
  return om_nccl_emu::depth() ? ncclSuccess : om_nccl_emu::flush();
}
