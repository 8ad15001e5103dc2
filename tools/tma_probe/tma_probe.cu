// Stand-alone probe for the TMA tile-load constraints the FAST kernel depends on (uint8 tensors, arbitrary x origin).
// Usage: tma_probe <rank> <boxW> <boxH> <x> <y> <pitch> <rows> <imgs> [paramPad]
// Prints OK + a checksum comparison, or the CUDA error.  Each variant runs in its own process.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct Pad { char b[2688]; };

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap m, int x, int y, int z, int bytes, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(sm + 64), dst = (unsigned)__cvta_generic_to_shared(sm + 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(dst), "l"(&m), "r"(x), "r"(y), "r"(mbar) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(dst), "l"(&m), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
  }
  __syncthreads();
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mbar) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[128 + i];
}

template <int RANK>
__global__ void probe_padded(const __grid_constant__ Pad pad, const __grid_constant__ CUtensorMap m, int x, int y, int z,
                             int bytes, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(sm + 64), dst = (unsigned)__cvta_generic_to_shared(sm + 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(&m), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
  }
  __syncthreads();
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mbar) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[128 + i] + pad.b[0];
}

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(e) do { cudaError_t r = (e); if (r != cudaSuccess) { printf("FAIL %s -> %s\n", #e, cudaGetErrorString(r)); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 9) return 2;
  const int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]);
  const int pitch = atoi(argv[6]), rows = atoi(argv[7]), imgs = atoi(argv[8]);
  const int padded = argc > 9 ? atoi(argv[9]) : 0;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaFree(0));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  enc_fn enc = (enc_fn)p;
  const size_t n = (size_t)pitch * rows * imgs;
  std::vector<uint8_t> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t *d, *dout;
  CK(cudaMalloc(&d, n));
  CK(cudaMalloc(&dout, 65536));
  CK(cudaMemcpy(d, h.data(), n, cudaMemcpyHostToDevice));
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)imgs};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("ENCODE_FAIL %d\n", (int)r); return 1; }
  const int bytes = bw * bh;
  const int z = imgs - 1;
  if (padded) {
    Pad pad{};
    probe_padded<3><<<1, 128, 128 + bytes>>>(pad, m, x, y, z, bytes, dout);
  } else if (rank == 2) {
    probe<2><<<1, 128, 128 + bytes>>>(m, x, y, 0, bytes, dout);
  } else {
    probe<3><<<1, 128, 128 + bytes>>>(m, x, y, z, bytes, dout);
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<uint8_t> o(bytes);
  CK(cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int j = 0; j < bh; ++j)
    for (int i = 0; i < bw; ++i) {
      const int gx = x + i, gy = y + j;
      uint8_t want = 0;
      if (gx >= 0 && gx < pitch && gy >= 0 && gy < rows)
        want = h[(size_t)(rank == 2 ? 0 : z) * pitch * rows + (size_t)gy * pitch + gx];
      bad += o[j * bw + i] != want;
    }
  printf("OK mismatches=%d\n", bad);
  return 0;
}
