// TEST INFRASTRUCTURE -- csrc/prune_csr.cu (K1) compiled for the host (see cuda_runtime.h in this directory); exports
// gpt_prune_csr taking HOST pointers.  Built by tests/emu/emu_build.py into the same library as deprel_host.cpp.
#define GPT_HOST_EMULATION 1
#include "cuda_runtime.h"

namespace {
thread_local int smem[200 * 1024 / 4];   // the kernel's `extern __shared__ int smem[]`: one block runs at a time
}

#include "../../gcn_over_pruned_trees_b200/csrc/prune_csr.cu"
