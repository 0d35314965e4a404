// synth.cuh -- counter-based synthetic embedding generator, device + host.
// Same definition as oracle/synth.py and vo_synth_rows (oracle/vm_oracle.c); kept in the
// product library only so that 1M..100M-row bench stores can be generated directly in HBM.
#pragma once
#include <stdint.h>

namespace vm {
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ int synth_int(uint64_t h, uint32_t byte)
{
    uint32_t b = (uint32_t)((h >> (8 * byte)) & 0xFF);
    return (int)((b * 255u) >> 8) - 127;
}
}  // namespace vm
