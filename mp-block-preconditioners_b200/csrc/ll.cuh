// LL elements: an 8-byte value and its 8-byte sequence tag written with ONE 16-byte store (the idea of NCCL's LL
// protocol).  A reader that sees the expected tag has the value: peer-memory exchanges built on these need no memory
// fence and no separate flag.  Used by the halo exchange (stencil.cuh) and the fused all-reduce (blas1.cuh).
#pragma once
#include <stdint.h>

namespace mpbp {

struct alignas(16) LLElem {
  double v;
  unsigned long long tag;
};
#ifdef MPBP_EMU
// host emulation (tests/emu): value first, then the tag with release / acquire ordering
__device__ __forceinline__ void st_ll(LLElem* p, double v, unsigned long long tag) {
  p->v = v;
  __atomic_store_n(&p->tag, tag, __ATOMIC_RELEASE);
}
__device__ __forceinline__ double ld_ll(const LLElem* p, unsigned long long& tag) {
  tag = __atomic_load_n(&p->tag, __ATOMIC_ACQUIRE);
  return p->v;
}
#else
__device__ __forceinline__ void st_ll(LLElem* p, double v, unsigned long long tag) {
  asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(tag));
}
__device__ __forceinline__ double ld_ll(const LLElem* p, unsigned long long& tag) {
  long long a;
  asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(tag) : "l"(p));
  return __longlong_as_double(a);
}
#endif
__device__ __forceinline__ double ll_wait(const LLElem* p, unsigned long long seq) {
  unsigned long long t;
  double v;
  do {
    v = ld_ll(p, t);
  } while (t != seq);
  return v;
}
// the same for a whole warp (all 32 lanes call it, each with its own element): the loop is warp-uniform, which keeps
// the code after it on the uniform datapath (a per-lane spin loop in front of the marching loops cost them spills)
__device__ __forceinline__ double ll_wait_warp(const LLElem* p, unsigned long long seq) {
#ifdef MPBP_EMU
  return ll_wait(p, seq);
#else
  unsigned long long t;
  double v;
  do {
    v = ld_ll(p, t);
  } while (!__all_sync(0xffffffffu, t == seq));
  return v;
#endif
}

}  // namespace mpbp
