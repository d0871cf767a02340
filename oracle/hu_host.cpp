// Host build of nellie_b200/csrc/hu.cu through oracle/cuda_emu.h (TEST INFRASTRUCTURE ONLY): the same nb200_hu_* entry
// points executing the same kernel bodies serially.  Built by __graft_entry__.build() with
//   g++ -O2 -ffp-contract=off -mfma -shared -fPIC -DNB200_HOST_EMU='"<repo>/oracle/cuda_emu.h"' -x c++ oracle/hu_host.cpp
// (-mfma: devmath.cuh's explicit fmaf() must be a real fused multiply-add).  Never loaded by nellie_b200.
#include "../nellie_b200/csrc/hu.cu"
