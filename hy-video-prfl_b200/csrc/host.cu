// Host-side plumbing: error strings, device check, TMA descriptor encoding, launch counter.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace prfl {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return PRFL_E_CUDA;
}

static std::atomic<int> g_dev_state[64];  // 0 unknown, 1 ok, -1 wrong arch (written from the forward and the autograd threads)
static std::atomic<int> g_sm_count[64];

static int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  return dev;
}
bool device_needs_init(unsigned long long* mask_storage) {
  auto* m = reinterpret_cast<std::atomic<unsigned long long>*>(mask_storage);
  return ((m->load(std::memory_order_acquire) >> current_device_slot()) & 1ull) == 0;
}
void device_mark_init(unsigned long long* mask_storage) {
  auto* m = reinterpret_cast<std::atomic<unsigned long long>*>(mask_storage);
  m->fetch_or(1ull << current_device_slot(), std::memory_order_release);
}

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (prfl_b200 has no CPU fallback)", cudaGetErrorString(e));
    cudaGetLastError();
    return PRFL_E_ARCH;
  }
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_dev_state[dev].load(std::memory_order_acquire) == 0) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    g_sm_count[dev].store(p.multiProcessorCount, std::memory_order_relaxed);
    g_dev_state[dev].store((p.major == 10) ? 1 : -1, std::memory_order_release);
    if (p.major != 10) set_error("device %d is sm_%d%d; prfl_b200 is sm_100a only", dev, p.major, p.minor);
  }
  if (g_dev_state[dev].load(std::memory_order_acquire) < 0) {
    set_error("device %d is not sm_100; prfl_b200 is sm_100a only (no fallback)", dev);
    return PRFL_E_ARCH;
  }
  return PRFL_OK;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  const int n = g_sm_count[dev].load(std::memory_order_relaxed);
  return n > 0 ? n : 148;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
  });
  return g_encode;
}

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, int swizzle128, int elem_bytes) {
  EncodeTiledFn f = get_encode();
  if (!f) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return PRFL_E_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base pointer not 16-byte aligned");
    return PRFL_E_ALIGN;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = f(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank %d dims %llu %llu %llu box %u %u %u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return PRFL_E_CUDA;
  }
  return PRFL_OK;
}

int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t stride1_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle128, int elem_bytes) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  if (stride1_bytes % 16 != 0) {
    set_error("TMA stride %llu not a multiple of 16 bytes", (unsigned long long)stride1_bytes);
    return PRFL_E_ALIGN;
  }
  return encode(m, base, 2, dims, strides, box, swizzle128, elem_bytes);
}

int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                 uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle128, int elem_bytes) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  if (stride1_bytes % 16 != 0 || stride2_bytes % 16 != 0) {
    set_error("TMA strides %llu/%llu not multiples of 16 bytes", (unsigned long long)stride1_bytes,
              (unsigned long long)stride2_bytes);
    return PRFL_E_ALIGN;
  }
  return encode(m, base, 3, dims, strides, box, swizzle128, elem_bytes);
}

}  // namespace prfl

extern "C" {
int prfl_abi_version(void) { return 2; }
const char* prfl_last_error_string(void) { return prfl::g_err; }
int64_t prfl_launch_count(void) { return prfl::g_launches.load(); }
void prfl_launch_count_reset(void) { prfl::g_launches.store(0); }
}
