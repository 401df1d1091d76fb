// Microbenchmark: streaming blend h = o*(1-m) + h*m over 36.1M floats, kernel variants.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ldg_nc(const float4* p) { float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ float4 ld_na(const float4* p) { float4 r; asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st_na(float4* p, float4 v) { asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld8(const void* p, int hint) {
  f8 r; uint32_t* u = reinterpret_cast<uint32_t*>(r.v);
  if (hint) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(u[0]),"=r"(u[1]),"=r"(u[2]),"=r"(u[3]),"=r"(u[4]),"=r"(u[5]),"=r"(u[6]),"=r"(u[7]) : "l"(p));
  else asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(u[0]),"=r"(u[1]),"=r"(u[2]),"=r"(u[3]),"=r"(u[4]),"=r"(u[5]),"=r"(u[6]),"=r"(u[7]) : "l"(p));
  return r; }
__device__ __forceinline__ void st8(void* p, const f8& r) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(r.v);
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "r"(u[0]),"r"(u[1]),"r"(u[2]),"r"(u[3]),"r"(u[4]),"r"(u[5]),"r"(u[6]),"r"(u[7]) : "memory"); }
__device__ __forceinline__ float4 ld_ef(const float4* p) { return ld_na(p); }
// 256-bit variant: one CTA per chunk of T*8*U floats
template <int T, int U, int HINT>
__global__ void __launch_bounds__(T) k_chunk8(const float* __restrict__ o, float* __restrict__ h, float m, float om) {
  const size_t base = ((size_t)blockIdx.x * T * U + threadIdx.x) * 8;
  f8 a[U], b[U];
#pragma unroll
  for (int u = 0; u < U; ++u) a[u] = ld8(o + base + (size_t)u * T * 8, HINT);
#pragma unroll
  for (int u = 0; u < U; ++u) b[u] = ld8(h + base + (size_t)u * T * 8, HINT);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    f8 r;
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = __fadd_rn(__fmul_rn(a[u].v[e], om), __fmul_rn(b[u].v[e], m));
    st8(h + base + (size_t)u * T * 8, r);
  }
}
__device__ __forceinline__ void st_cs(float4* p, float4 v) { asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
__device__ __forceinline__ float bl(float o, float h, float m, float om) { return __fadd_rn(__fmul_rn(o, om), __fmul_rn(h, m)); }
__device__ __forceinline__ float4 bl4(float4 o, float4 h, float m, float om) { return make_float4(bl(o.x,h.x,m,om), bl(o.y,h.y,m,om), bl(o.z,h.z,m,om), bl(o.w,h.w,m,om)); }

// chunked: one CTA per chunk of T*4*U floats
template <int T, int U, int HINT>
__global__ void __launch_bounds__(T) k_chunk(const float4* __restrict__ o, float4* __restrict__ h, float m, float om) {
  const size_t base = (size_t)blockIdx.x * T * U + threadIdx.x;
  float4 a[U], b[U];
#pragma unroll
  for (int u = 0; u < U; ++u) a[u] = HINT ? ld_ef(o + base + u * T) : ldg_nc(o + base + u * T);
#pragma unroll
  for (int u = 0; u < U; ++u) b[u] = HINT ? ld_ef(h + base + u * T) : ld_na(h + base + u * T);
#pragma unroll
  for (int u = 0; u < U; ++u) { if (HINT) st_cs(h + base + u * T, bl4(a[u], b[u], m, om)); else st_na(h + base + u * T, bl4(a[u], b[u], m, om)); }
}
// persistent grid-stride
template <int T, int U>
__global__ void __launch_bounds__(T) k_persist(const float4* __restrict__ o, float4* __restrict__ h, size_t n4, float m, float om) {
  for (size_t base = (size_t)blockIdx.x * T * U + threadIdx.x; base + (U - 1) * T < n4; base += (size_t)gridDim.x * T * U) {
    float4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = ldg_nc(o + base + u * T);
#pragma unroll
    for (int u = 0; u < U; ++u) b[u] = ld_na(h + base + u * T);
#pragma unroll
    for (int u = 0; u < U; ++u) st_na(h + base + u * T, bl4(a[u], b[u], m, om));
  }
}
__global__ void k_copy(const float4* __restrict__ o, float4* __restrict__ h, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) h[i] = o[i];
}

template <class F> float timeit(F f, int reps = 20) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(e0); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps * 1e3f;
}
static void* g_flush = nullptr;
template <class F> float timeit_cold(F f, int reps = 10) {
  if (!g_flush) cudaMalloc(&g_flush, 512u << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float tot = 0;
  for (int i = 0; i < reps; ++i) {
    cudaMemsetAsync(g_flush, i, 512u << 20);
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms;
  }
  return tot / reps * 1e3f;
}
int main() {
  const size_t n = 36095168; const size_t n4 = n / 4;
  float *o, *h; cudaMalloc(&o, n * 4); cudaMalloc(&h, n * 4); cudaMemset(o, 0, n * 4); cudaMemset(h, 0, n * 4);
  const double bytes = 12.0 * n;
  auto rep = [&](const char* name, float us) { printf("%-34s %7.1f us  %7.0f GB/s\n", name, us, bytes / us / 1e3); };
#define CH(T, U, H) rep("chunk T=" #T " U=" #U " hint=" #H, timeit([&] { k_chunk<T, U, H><<<(unsigned)(n4 / (T * U)), T>>>((const float4*)o, (float4*)h, 0.999f, 0.001f); }))
  CH(256, 4, 0); CH(256, 8, 0); CH(512, 4, 0); CH(128, 8, 0); CH(256, 2, 0); CH(1024, 2, 0); CH(256, 4, 1); CH(256, 8, 1); CH(512, 4, 1);
#define C8(T, U, H) rep("chunk256b T=" #T " U=" #U " hint=" #H, timeit([&] { k_chunk8<T, U, H><<<(unsigned)(n / (T * U * 8)), T>>>(o, h, 0.999f, 0.001f); }))
  C8(256, 2, 0); C8(256, 4, 0); C8(128, 4, 0); C8(512, 2, 0); C8(256, 2, 1); C8(256, 4, 1);
#define PE(T, U, G) rep("persist T=" #T " U=" #U " grid=" #G, timeit([&] { k_persist<T, U><<<G, T>>>((const float4*)o, (float4*)h, n4, 0.999f, 0.001f); }))
  PE(256, 4, 148 * 8); PE(256, 8, 148 * 4); PE(512, 4, 148 * 4); PE(256, 4, 148 * 16); PE(1024, 2, 148 * 2);
  rep("COLD chunk T=256 U=4", timeit_cold([&] { k_chunk<256, 4, 0><<<(unsigned)(n4 / (256 * 4)), 256>>>((const float4*)o, (float4*)h, 0.999f, 0.001f); }));
  rep("COLD chunk T=1024 U=2", timeit_cold([&] { k_chunk<1024, 2, 0><<<(unsigned)(n4 / (1024 * 2)), 1024>>>((const float4*)o, (float4*)h, 0.999f, 0.001f); }));
  rep("COLD chunk256b T=256 U=4", timeit_cold([&] { k_chunk8<256, 4, 0><<<(unsigned)(n / (256 * 4 * 8)), 256>>>(o, h, 0.999f, 0.001f); }));
  rep("COLD persist T=256 U=8 g=592", timeit_cold([&] { k_persist<256, 8><<<148 * 4, 256>>>((const float4*)o, (float4*)h, n4, 0.999f, 0.001f); }));
  { float usc = timeit_cold([&] { k_copy<<<148 * 16, 512>>>((const float4*)o, (float4*)h, n4); }); printf("%-34s %7.1f us  %7.0f GB/s (r+w)\n", "COLD copy grid-stride", usc, 8.0 * n / usc / 1e3); }
  float us = timeit([&] { k_copy<<<148 * 16, 512>>>((const float4*)o, (float4*)h, n4); });
  printf("%-34s %7.1f us  %7.0f GB/s (r+w)\n", "copy grid-stride", us, 8.0 * n / us / 1e3);
  us = timeit([&] { cudaMemcpyAsync(h, o, n * 4, cudaMemcpyDeviceToDevice); });
  printf("%-34s %7.1f us  %7.0f GB/s (r+w)\n", "cudaMemcpy D2D 144MB", us, 8.0 * n / us / 1e3);
  return 0;
}
