// orbx_api.cu — context, error text, version.  There is deliberately no CPU fallback: without a
// usable CUDA device orbx_create() fails.
#include "orbx_common.cuh"
#include <cstdarg>

static thread_local std::string g_err;

void orbx_set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

extern "C" {

int orbx_abi_version(void) { return 1; }
const char* orbx_last_error(void) { return g_err.c_str(); }

orbx_ctx* orbx_create(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    orbx_set_error("orbx_create: no CUDA device (%s); liborbx has no CPU path",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    orbx_set_error("orbx_create: device %d out of range [0,%d)", device, n);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    orbx_set_error("orbx_create: cudaSetDevice(%d) failed", device);
    return nullptr;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major < 10) {
    orbx_set_error("orbx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                   prop.major, prop.minor);
    return nullptr;
  }
  orbx_ctx* c = new orbx_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    orbx_set_error("orbx_create: cudaStreamCreate failed");
    delete c;
    return nullptr;
  }
  return c;
}

void orbx_destroy(orbx_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->arenaDev) cudaFree(c->arenaDev);
  if (c->arenaHost) cudaFreeHost(c->arenaHost);
  delete c;
}

void* orbx_stream(orbx_ctx* c) { return c ? (void*)c->stream : nullptr; }

int orbx_synchronize(orbx_ctx* c) {
  if (!c) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(c->device));
  ORBX_CUDA(cudaDeviceSynchronize());
  return ORBX_OK;
}

uint64_t orbx_launch_count(const orbx_ctx* c) { return c ? c->launches.load() : 0; }

void* orbx_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    orbx_set_error("orbx_host_alloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
void orbx_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
