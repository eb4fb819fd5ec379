// TEST INFRASTRUCTURE -- csrc/aggregate.cu (K2) compiled for the host (see cuda_runtime.h, emu_smem_ops.h in this
// directory); exports gpt_gcn_aggregate_{fwd,bwd,bwd_pre,fwd_pool,bwd_pool} taking HOST pointers.  The cp.async path
// of the kernels runs (no tensor maps on the host).  Built by tests/emu/emu_build.py.
#define GPT_HOST_EMULATION 1
#include "cuda_runtime.h"
#include "emu_smem_ops.h"

namespace {
// the kernels' `extern __shared__ __align__(128) unsigned char smem_raw[]`: one block runs at a time
thread_local __attribute__((aligned(128))) unsigned char smem_raw[224 * 1024];
}

#include "../../gcn_over_pruned_trees_b200/csrc/aggregate.cu"
