// pairs.cu -- all-pairs threshold scorer (placeholder until the kernel lands in this round).
#include "common.cuh"
namespace vm {
int k_pairs_above(int, const void *, int, int64_t, int, int, float, int64_t, int64_t *, int64_t *, float *, int64_t *, int, int,
                  int, cudaStream_t)
{
    set_error("pairs kernel not built");
    return VM_ERR_UNSUPPORTED;
}
}  // namespace vm
