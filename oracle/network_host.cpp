// Host build of nellie_b200/csrc/network.cu through oracle/cuda_emu.h (TEST INFRASTRUCTURE ONLY): the same
// nb200_network_* entry points executing the same kernel bodies serially.  Built by __graft_entry__.build() with
//   g++ -O2 -ffp-contract=off -mfma -shared -fPIC -DNB200_HOST_EMU='"<repo>/oracle/cuda_emu.h"' -x c++ oracle/network_host.cpp
// Never loaded by nellie_b200.
#include "../nellie_b200/csrc/network.cu"
