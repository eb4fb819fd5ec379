// TEST INFRASTRUCTURE -- csrc/embed.cu (K5) compiled for the host (see cuda_runtime.h in this directory); exports
// gpt_embed_fwd / gpt_embed_bwd / gpt_embed_rows_sqnorm / gpt_embed_rows_sgd taking HOST pointers.  Built by
// tests/emu/emu_build.py.
#define GPT_HOST_EMULATION 1
#include "cuda_runtime.h"

namespace {
thread_local float s_small[48 * 1024 / 4];   // the backward kernel's `extern __shared__ float s_small[]`
}

#include "../../gcn_over_pruned_trees_b200/csrc/embed.cu"
