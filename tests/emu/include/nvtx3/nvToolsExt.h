// TEST INFRASTRUCTURE - NVTX stand-in of the CPU SIMT emulation build (tests/emu): ranges are no-ops.
#pragma once
static inline int nvtxRangePushA(const char*) { return 0; }
static inline int nvtxRangePop() { return 0; }
