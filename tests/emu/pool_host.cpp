// TEST INFRASTRUCTURE -- csrc/pool3.cu (K4) compiled for the host (see cuda_runtime.h in this directory); exports
// gpt_pool3_fwd / gpt_pool3_bwd / gpt_pool3_bwd_masked taking HOST pointers.  Built by tests/emu/emu_build.py.
#define GPT_HOST_EMULATION 1
#include "cuda_runtime.h"

namespace {
// the kernels' dynamic shared memory (`extern __shared__ ...`): one block runs at a time
thread_local __attribute__((aligned(16))) unsigned char smem_raw[200 * 1024];
thread_local unsigned char s_flags[48 * 1024];
}

#include "../../gcn_over_pruned_trees_b200/csrc/pool3.cu"
