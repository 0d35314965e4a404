// scan_tc.cu -- tcgen05/TMA streaming scorer (placeholder until the kernel lands in this round).
#include "common.cuh"
namespace vm {
bool scan_tc_supported(int, int, int, int) { return false; }
int launch_scan_tc(const ScanArgs &, const void *) { set_error("tcgen05 scan not built"); return VM_ERR_UNSUPPORTED; }
}  // namespace vm
