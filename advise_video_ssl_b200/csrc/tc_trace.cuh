// Developer-only timeline probe (tools/microbench/tc_trace.cu builds with AVSSL_TC_TRACE):
// per-role clock64 stamps of CTA (0,0).  Compiles to nothing in the product build.
#pragma once
namespace avssl {
#ifdef AVSSL_TC_TRACE
// developer build only (tools/microbench/tc_trace.cu): per-role timestamps of CTA (0,0)
__device__ long long g_tc_trace[16][64];
__device__ int g_tc_probe;  // experiment selector of the trace build
#define TC_PROBE(x) (g_tc_probe == (x))
#define TC_TRACE(ev, t)                                                                \
  do {                                                                                 \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (t) < 64) g_tc_trace[ev][t] = clock64(); \
  } while (0)
#else
#define TC_TRACE(ev, t) \
  do {                  \
  } while (0)
#define TC_PROBE(x) false
#endif

}  // namespace avssl
