// TEST INFRASTRUCTURE — host emulation build of the CUDA core sources (rustracer_b200/csrc/*.h with -DRT_EMU).
// Launches become loops and atomics become plain read-modify-writes, so the BVH builder, the traversal state
// machine and the wavefront shading code can be compared against the oracle on the GPU-less build box.
// Exports carry an emu_ prefix; the rustracer_b200 package never loads this library and the product library
// (librt_b200.so) contains none of it.  GPU parity is established separately by the `-m gpu` tests.
#define RT_EMU 1
#include "../../rustracer_b200/csrc/rt_core.h"
