// placeholder until the backward kernel lands
#include "common.cuh"
#include "scan_params.h"
namespace vmb {
int scan_bwd_launch(const ScanBwdParams&, int, cudaStream_t) {
    set_error("selective_scan_bwd: not built yet");
    return VMB_ERR_INVALID;
}
}  // namespace vmb
