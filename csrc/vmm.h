// Symmetric memory for the sharded parameter server: CUDA virtual-memory-management allocations that every GPU of
// the box maps (peer access over NVLink), and NVSwitch multicast objects (NVLS) whose alias address replicates one
// `multimem.st` into every bound GPU's memory.
//
//   physical allocation : cuMemCreate (POSIX-fd shareable) on the owning device
//   mapping             : cuMemAddressReserve + cuMemMap + cuMemSetAccess(read/write for the mapping device)
//   cross-process       : cuMemExportToShareableHandle -> fd -> (SCM_RIGHTS, done in Python) -> cuMemImportFromShareableHandle
//   multicast           : cuMulticastCreate(N devices) -> every process cuMulticastAddDevice(own device) -> barrier ->
//                         cuMulticastBindMem(own physical allocation) -> map the multicast handle like any allocation
//
// The driver API is reached through cudaGetDriverEntryPoint (the library links only cudart statically + libdl).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <vector>

namespace sfvmm {

template <class Fn>
inline Fn drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr)
    throw std::runtime_error(std::string("sparkflow_b200: driver entry point not available: ") + name);
  return reinterpret_cast<Fn>(p);
}

inline void ckd(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) {
    const char* msg = nullptr;
    static auto get = drv<CUresult (*)(CUresult, const char**)>("cuGetErrorString");
    get(r, &msg);
    throw std::runtime_error(std::string("sparkflow_b200 driver error in ") + what + ": " + (msg ? msg : "?") + " (" + std::to_string(static_cast<int>(r)) + ")");
  }
}

inline CUmemAllocationProp prop_for(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

inline size_t granularity(int device, bool multicast, int n_devices) {
  size_t g = 0;
  const CUmemAllocationProp p = prop_for(device);
  static auto fn = drv<CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags)>("cuMemGetAllocationGranularity");
  ckd(fn(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  if (multicast) {
    CUmulticastObjectProp mp{};
    mp.numDevices = static_cast<unsigned int>(n_devices);
    mp.size = g;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    static auto mfn = drv<CUresult (*)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags)>("cuMulticastGetGranularity");
    ckd(mfn(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
    if (mg > g) g = mg;
  }
  return g;
}

inline bool multicast_supported(int device) {
  int v = 0;
  static auto fn = drv<CUresult (*)(int*, CUdevice_attribute, CUdevice)>("cuDeviceGetAttribute");
  if (fn(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device) != CUDA_SUCCESS) return false;
  return v != 0;
}

// physical allocation on `device` (bytes must be a multiple of the granularity)
inline uint64_t create(int device, size_t bytes) {
  CUmemGenericAllocationHandle h = 0;
  const CUmemAllocationProp p = prop_for(device);
  static auto fn = drv<CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long)>("cuMemCreate");
  ckd(fn(&h, bytes, &p, 0), "cuMemCreate");
  return static_cast<uint64_t>(h);
}
inline void release(uint64_t h) {
  static auto fn = drv<CUresult (*)(CUmemGenericAllocationHandle)>("cuMemRelease");
  fn(static_cast<CUmemGenericAllocationHandle>(h));
}
inline int export_fd(uint64_t h) {
  int fd = -1;
  static auto fn = drv<CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long)>("cuMemExportToShareableHandle");
  ckd(fn(&fd, static_cast<CUmemGenericAllocationHandle>(h), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
  return fd;
}
inline uint64_t import_fd(int fd) {
  CUmemGenericAllocationHandle h = 0;
  static auto fn = drv<CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType)>("cuMemImportFromShareableHandle");
  ckd(fn(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle");
  return static_cast<uint64_t>(h);
}
// map `h` (a physical allocation or a multicast object) into this process and give `devices` read/write access
inline uint64_t map(uint64_t h, size_t bytes, const std::vector<int>& devices) {
  CUdeviceptr va = 0;
  static auto reserve = drv<CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long)>("cuMemAddressReserve");
  static auto mmap = drv<CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long)>("cuMemMap");
  static auto access = drv<CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t)>("cuMemSetAccess");
  ckd(reserve(&va, bytes, 0, 0, 0), "cuMemAddressReserve");
  ckd(mmap(va, bytes, 0, static_cast<CUmemGenericAllocationHandle>(h), 0), "cuMemMap");
  std::vector<CUmemAccessDesc> acc(devices.size());
  for (size_t i = 0; i < devices.size(); ++i) {
    acc[i].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[i].location.id = devices[i];
    acc[i].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  ckd(access(va, bytes, acc.data(), acc.size()), "cuMemSetAccess");
  return static_cast<uint64_t>(va);
}
inline void unmap(uint64_t va, size_t bytes) {
  static auto unmap_fn = drv<CUresult (*)(CUdeviceptr, size_t)>("cuMemUnmap");
  static auto free_fn = drv<CUresult (*)(CUdeviceptr, size_t)>("cuMemAddressFree");
  unmap_fn(static_cast<CUdeviceptr>(va), bytes);
  free_fn(static_cast<CUdeviceptr>(va), bytes);
}

// ---- multicast (NVLS) ----
inline uint64_t mc_create(int n_devices, size_t bytes) {
  CUmulticastObjectProp mp{};
  mp.numDevices = static_cast<unsigned int>(n_devices);
  mp.size = bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h = 0;
  static auto fn = drv<CUresult (*)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*)>("cuMulticastCreate");
  ckd(fn(&h, &mp), "cuMulticastCreate");
  return static_cast<uint64_t>(h);
}
inline void mc_add_device(uint64_t mc, int device) {
  static auto fn = drv<CUresult (*)(CUmemGenericAllocationHandle, CUdevice)>("cuMulticastAddDevice");
  ckd(fn(static_cast<CUmemGenericAllocationHandle>(mc), device), "cuMulticastAddDevice");
}
inline void mc_bind(uint64_t mc, size_t mc_offset, uint64_t mem, size_t mem_offset, size_t bytes) {
  static auto fn = drv<CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long)>("cuMulticastBindMem");
  ckd(fn(static_cast<CUmemGenericAllocationHandle>(mc), mc_offset, static_cast<CUmemGenericAllocationHandle>(mem), mem_offset, bytes, 0), "cuMulticastBindMem");
}
inline void mc_unbind(uint64_t mc, int device, size_t mc_offset, size_t bytes) {
  static auto fn = drv<CUresult (*)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t)>("cuMulticastUnbind");
  fn(static_cast<CUmemGenericAllocationHandle>(mc), device, mc_offset, bytes);
}

}  // namespace sfvmm
