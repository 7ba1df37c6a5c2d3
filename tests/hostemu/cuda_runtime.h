// Stub <cuda_runtime.h> of the host emulation (tests/hostemu): see cuda_emu.h.  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"
