// Host-side helpers shared by the translation units of liboron_b200.so.
#pragma once
#include <atomic>
#include <cstdint>

#include <cuda_runtime.h>

struct CUtensorMap_st;

namespace oron {
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
int num_sms();
bool pdl_enabled();
// bf16 (any 16-bit type) tensor [d2][d1][d0] (d0 contiguous), SWIZZLE_128B, box = (64, box1, 1); OOB reads give zeros.
// `m` is a CUtensorMap* (kept opaque here so that this header does not need cuda.h).
int make_tmap_bf16(::CUtensorMap_st* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                   uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box1, int rank);

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel's prologue (barrier init, TMEM
// allocation, descriptor prefetch) overlaps the tail of the previous kernel in the stream; every kernel launched
// this way executes griddepcontrol.wait before reading upstream results. ORON_PDL=0 disables it.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
}  // namespace oron
