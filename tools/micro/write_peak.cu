// Write-only HBM bandwidth on B200 for the store patterns of the cut kernels (diagnostic, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_peak write_peak.cu && ./write_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// contiguous: every thread writes VEC bytes, consecutive threads consecutive addresses
template <int VEC>
__global__ void k_fill_contig(uint8_t* p, size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
  for (; i + VEC <= n; i += stride) {
    if (VEC == 32) {
      asm volatile("st.global.v4.b64 [%0], {%1,%1,%1,%1};" ::"l"(p + i), "l"(0x0101010101010101ull) : "memory");
    } else if (VEC == 16) {
      *reinterpret_cast<uint4*>(p + i) = make_uint4(1, 1, 1, 1);
    } else {
      *reinterpret_cast<uint2*>(p + i) = make_uint2(1, 1);
    }
  }
}

// classify-like: a warp owns runs of RUN bytes (lane = 96 contiguous bytes = 3 x 32-byte stores) in 8 rows `rowstride` apart,
// marching through layers `laystride` apart
__global__ void k_fill_tiles(uint8_t* p, int ntx, int nty, int nlay, int zs, size_t rowstride, size_t laystride) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int tx = warp % ntx, ty = (warp / ntx) % nty, tz = warp / ((size_t)ntx * nty);
  if (tz * zs >= nlay) return;
  const int task = lane / 2, sub = lane & 1;
  const int wwc = task % 2, rc = task / 2;
  uint8_t* out = p + (size_t)(tz * zs) * laystride + (size_t)(ty * 8 + rc) * rowstride + (size_t)(tx * 2 + wwc) * 192 + sub * 96;
  for (int k = 0; k < zs && tz * zs + k < nlay; ++k, out += laystride) {
#pragma unroll
    for (int q = 0; q < 3; ++q)
      asm volatile("st.global.v4.b64 [%0], {%1,%1,%1,%1};" ::"l"(out + 32 * q), "l"(0x0101010101010101ull) : "memory");
  }
}

// row-major tiles: a warp writes one whole 3072-byte row (lane = 96 bytes), 8 rows, marching layers
__global__ void k_fill_rows(uint8_t* p, int nrow_tiles, int nlay, int zs, size_t rowstride, size_t laystride, int order) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int ty = warp % nrow_tiles, tz = warp / nrow_tiles;
  if (tz * zs >= nlay) return;
  for (int k = 0; k < zs && tz * zs + k < nlay; ++k) {
    for (int r = 0; r < 8; ++r) {
      uint8_t* out = p + (size_t)(tz * zs + k) * laystride + (size_t)(ty * 8 + r) * rowstride;
      if (order == 0) {  // lane-contiguous 96 bytes
#pragma unroll
        for (int q = 0; q < 3; ++q)
          asm volatile("st.global.v4.b64 [%0], {%1,%1,%1,%1};" ::"l"(out + lane * 96 + 32 * q), "l"(0x0101010101010101ull) : "memory");
      } else {  // instruction-contiguous: 1024 bytes per instruction
#pragma unroll
        for (int q = 0; q < 3; ++q)
          asm volatile("st.global.v4.b64 [%0], {%1,%1,%1,%1};" ::"l"(out + q * 1024 + lane * 32), "l"(0x0101010101010101ull) : "memory");
      }
    }
  }
}

__global__ void k_copy(const uint4* a, uint4* b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}

template <class F>
float timeit(F f, int reps = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  f();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t N = (size_t)512 * 512 * 512 * 6;  // the tet state row of the 512^3 model: 805 MB
  uint8_t *p, *q;
  cudaMalloc(&p, N);
  cudaMalloc(&q, N);
  auto rep = [&](const char* name, float ms, double bytes) { printf("%-58s %8.4f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
  rep("cudaMemsetAsync 805 MB", timeit([&] { cudaMemsetAsync(p, 1, N); }), (double)N);
  for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    char nm[128];
    snprintf(nm, sizeof nm, "contiguous st.v4.b64 (32 B/thread), %d blocks x 256", blocks);
    rep(nm, timeit([&] { k_fill_contig<32><<<blocks, 256>>>(p, N); }), (double)N);
  }
  rep("contiguous 16 B/thread, 148*16 x 256", timeit([&] { k_fill_contig<16><<<148 * 16, 256>>>(p, N); }), (double)N);
  rep("contiguous 8 B/thread, 148*16 x 256", timeit([&] { k_fill_contig<8><<<148 * 16, 256>>>(p, N); }), (double)N);
  {
    const int ntx = 8, nty = 64, nlay = 512;
    for (int zs : {8, 22, 64}) {
      const size_t ntile = (size_t)ntx * nty * ((nlay + zs - 1) / zs);
      char nm[128];
      snprintf(nm, sizeof nm, "classify tiles (384 B runs x 8 rows), zs=%d, 128 thr/blk", zs);
      rep(nm, timeit([&] { k_fill_tiles<<<(unsigned)((ntile + 3) / 4), 128>>>(p, ntx, nty, nlay, zs, 3072, (size_t)512 * 3072); }), (double)N);
    }
    for (int order : {0, 1})
      for (int zs : {8, 22}) {
        const size_t ntile = (size_t)64 * ((nlay + zs - 1) / zs);
        char nm[128];
        snprintf(nm, sizeof nm, "row tiles (3072 B runs x 8 rows), zs=%d, order=%d", zs, order);
        rep(nm, timeit([&] { k_fill_rows<<<(unsigned)((ntile + 3) / 4), 128>>>(p, 64, nlay, zs, 3072, (size_t)512 * 3072, order); }), (double)N);
      }
  }
  rep("copy 805 MB -> 805 MB (uint4), bytes = read + write", timeit([&] { k_copy<<<148 * 16, 256>>>((const uint4*)p, (uint4*)q, N / 16); }), 2.0 * N);
  rep("cudaMemcpyAsync D2D 805 MB, bytes = read + write", timeit([&] { cudaMemcpyAsync(q, p, N, cudaMemcpyDeviceToDevice); }), 2.0 * N);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
